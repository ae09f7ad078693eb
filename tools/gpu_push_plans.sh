#!/bin/bash
# tools/gpu_push_plans.sh <tag> <N> — weak bench on N GPUs, push mode vs deep halo
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
    print("$name", "ms", round(d["ms_per_step"],4), d["step_ms_slowest_rank"], "plans", [(p["temporal_block"],p["rows_per_warp"]) for p in d["slabs"]["tile_plans_per_rank"]][:3], "halo", d["slabs"]["halo_rows"], "push", d["slabs"]["push_mode"], "e2e", round(d["e2e"]["value"]/1e9,2))
except Exception as e:
    print("$name failed", e)
PY
}
run push SAYAL_TILE_ROWS=8
run deep118 SAYAL_SLAB_PUSH=0 SAYAL_SLAB_HALO=118
run deep118_rows10 SAYAL_SLAB_PUSH=0 SAYAL_SLAB_HALO=118 SAYAL_TILE_ROWS=10
run deep118_rows12 SAYAL_SLAB_PUSH=0 SAYAL_SLAB_HALO=118 SAYAL_TILE_ROWS=12
python bench.py --skip-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single ms', d['ms_per_step'], 'plan', d['plan']['temporal_block'], d['plan']['tile_rows_per_warp'], 'strong1', d['strong_16384'])"
