#!/bin/bash
# tools/gpu_bench_n.sh <tag> <N> — the driver's N-GPU bench line on one box (weak 1080 rows/GPU + parity vs one GPU +
# the 16384^2 strong record), stderr kept.
set -u
tag=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SAYAL_BENCH_DEBUG=1 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_n$N.json"))
print("ms", round(d["ms_per_step"],4), "value", round(d["value"]/1e9,2), d["step_ms_slowest_rank"], d["slabs"]["halo_rows"], d["slabs"]["push_mode"], d["slabs"].get("edge_slab_bonus_rows"), [ (p["temporal_block"],p["rows_per_warp"],p["local_rows"]) for p in d["slabs"]["tile_plans_per_rank"]])
print("parity", d["parity_vs_single_gpu"]); print("strong", d["strong_16384"]); print("e2e", d["e2e"])
PY
grep "rank" gpurun_out/${tag}_bench_n$N.err | tail -$N
