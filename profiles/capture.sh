#!/bin/bash
# profiles/capture.sh <tag> — run under gpurun on ONE B200.  Writes into gpurun_out/:
#   <tag>_launches.csv      every launch of `bench.py --steps 2 --warmup 3` with its device time
#   <tag>_projection.ncu-rep  ncu --set full of projection_pack_kernel (2 launches, fixed plan, no autotune)
#   <tag>_advect.ncu-rep      ncu --set full of advect_velocity / advect_smoke (1 launch each)
# Numbers printed by a run under ncu are never bench values.
set -u
tag=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:projection_pack -s 21 -c 2 -f \
    -o gpurun_out/${tag}_projection python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_projection.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:advect_ -s 6 -c 2 -f \
    -o gpurun_out/${tag}_advect python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/${tag}_advect.log 2>&1
ls -la gpurun_out/
