#!/bin/bash
set -u
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
{
for dbg in 0 1 2 3 4 7; do
echo "== marginals, push mode halo 18, SAYAL_DEBUG_PUSH=$dbg"; SAYAL_DEBUG_PUSH=$dbg timeout 300 $TR tools/slab_marginal.py --halos 18 2>/dev/null | grep "^N="
done
echo "== no PDL"; SAYAL_USE_PDL=0 timeout 300 $TR tools/slab_marginal.py --halos 18 2>/dev/null | grep "^N="
echo "== rows 12"; SAYAL_TILE_ROWS=12 timeout 300 $TR tools/slab_marginal.py --halos 18 2>/dev/null | grep "^N="
echo "== rows 12 dbg 7"; SAYAL_DEBUG_PUSH=7 SAYAL_TILE_ROWS=12 timeout 300 $TR tools/slab_marginal.py --halos 18 2>/dev/null | grep "^N="
} 2>&1 | tee gpurun_out/push_probe2_n$N.log
