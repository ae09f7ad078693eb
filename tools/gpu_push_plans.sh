#!/bin/bash
# tools/gpu_push_plans.sh <tag> <N> — weak bench on N GPUs under several slab settings
set -u
tag=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() {
  name=$1; shift
  env "$@" SAYAL_BENCH_SKIP_STRONG=1 SAYAL_BENCH_SKIP_PARITY=1 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_push_${name}_n$N.json 2> gpurun_out/${tag}_push_${name}_n$N.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_push_${name}_n$N.json"))
    print("$name", "ms", round(d["ms_per_step"],4), d["step_ms_slowest_rank"]["median"], "plans", [(p["temporal_block"],p["rows_per_warp"]) for p in d["slabs"]["tile_plans_per_rank"]][:3], "halo", d["slabs"]["halo_rows"], "push", d["slabs"]["push_mode"], "e2e", round(d["e2e"]["value"]/1e9,2))
except Exception as e:
    print("$name failed", e)
PY
}
run auto A=1
run x4 SAYAL_XCHG_BLOCKS=4
run x16 SAYAL_XCHG_BLOCKS=16
run x32 SAYAL_XCHG_BLOCKS=32
run x64 SAYAL_XCHG_BLOCKS=64
run nobalance SAYAL_SLAB_NO_BALANCE=1
