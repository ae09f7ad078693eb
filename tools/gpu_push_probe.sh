#!/bin/bash
set -u
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
{
echo "== marginals, deep halo 118"; SAYAL_SLAB_PUSH=0 timeout 300 $TR tools/slab_marginal.py --halos 118 2>/dev/null | grep "^N="
echo "== marginals, deep halo 118, rows 8 T 8 forced"; SAYAL_TILE_ROWS=8 SAYAL_TEMPORAL_BLOCK=8 SAYAL_SLAB_PUSH=0 timeout 300 $TR tools/slab_marginal.py --halos 118 2>/dev/null | grep "^N="
echo "== stage times (eager)"; timeout 300 $TR tools/slab_stage_times.py 2>/dev/null | grep "^rank"
} 2>&1 | tee gpurun_out/push_probe4_n$N.log
