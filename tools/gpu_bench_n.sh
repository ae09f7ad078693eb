#!/bin/bash
# tools/gpu_bench_n.sh <tag> <N> — the driver's N-GPU bench line on one box (weak 1080 rows/GPU + parity vs one GPU +
# the 16384^2 strong record), stderr kept; then the deep-halo schedule of round 1 for comparison (no parity / strong).
set -u
tag=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_slab_gpu.py -m gpu -x -q 2>&1 | tail -3
SAYAL_BENCH_DEBUG=1 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "rc=$?"; cat gpurun_out/${tag}_bench_n$N.json; grep "rank" gpurun_out/${tag}_bench_n$N.err | tail -$N
SAYAL_SLAB_PUSH=0 SAYAL_SLAB_HALO=118 SAYAL_BENCH_SKIP_PARITY=1 SAYAL_BENCH_SKIP_STRONG=1 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${N}_deephalo.json 2> gpurun_out/${tag}_bench_n${N}_deephalo.err
echo "deep halo rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_n${N}_deephalo.json
