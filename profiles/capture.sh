#!/bin/bash
# profiles/capture.sh <tag> — run under gpurun on ONE B200.  Writes into gpurun_out/:
#   <tag>_launches.csv      every launch of `bench.py --steps 2 --warmup 3` with its device time
#   <tag>_projection.ncu-rep  ncu --set full of projection_pack_kernel (2 launches of the plan the bench uses)
#   <tag>_advect.ncu-rep      ncu --set full of advect_velocity / advect_smoke (1 launch each)
# The plan (temporal block T, rows per warp) is read from an un-profiled bench run first and then pinned with
# SAYAL_AUTOTUNE=0 so that no tuning candidates appear in the launch list.  Numbers printed by a run under ncu
# are never bench values.
set -u
tag=${1:-r1}
mkdir -p gpurun_out
SAYAL_BENCH_SKIP_STRONG=1 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_plan.json 2>/dev/null
T=$(python -c "import json;print(json.load(open('gpurun_out/${tag}_plan.json'))['plan']['temporal_block'])")
R=$(python -c "import json;print(json.load(open('gpurun_out/${tag}_plan.json'))['plan']['tile_rows_per_warp'])")
echo "plan: T=$T rows=$R"
export SAYAL_BENCH_SKIP_STRONG=1 SAYAL_BENCH_NO_GATE=1
export SAYAL_AUTOTUNE=0 SAYAL_TEMPORAL_BLOCK=$T SAYAL_TILE_ROWS=$R
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:projection_pack -s 12 -c 2 -f \
    -o gpurun_out/${tag}_projection python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_projection.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:advect_ -s 6 -c 2 -f \
    -o gpurun_out/${tag}_advect python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_advect.log 2>&1
ls -la gpurun_out/
