#!/bin/bash
# tools/gpu_scaling.sh <tag> <N> [quick] — on a box with N GPUs: weak probe at 1080 rows/GPU over several halos, strong
# scaling of 16384 x 16384 (n = 50) at N and at 1 GPU of the same box, weak 16384 x 8192/GPU n = 200, bench.py --gpus N.
set -u
tag=${1:-r1}; N=${2:-2}; quick=${3:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
out=gpurun_out/${tag}_scaling_n$N.jsonl
: > $out
timeout 600 $TR --nproc-per-node $N tools/scaling.py --weak --width 1920 --height 1080 --halos 18,26,34,42,50,66,118 --steps 50 >> $out 2> gpurun_out/${tag}_scaling_n$N.err
timeout 900 $TR --nproc-per-node $N tools/scaling.py --width 16384 --height 16384 --halos 32,50,118 --steps 8 >> $out 2>> gpurun_out/${tag}_scaling_n$N.err
if [ -z "$quick" ]; then
  timeout 900 $TR --nproc-per-node 1 tools/scaling.py --width 16384 --height 16384 --steps 8 >> $out 2>> gpurun_out/${tag}_scaling_n$N.err
  timeout 900 $TR --nproc-per-node 1 tools/scaling.py --width 16384 --height 8192 --iters 200 --steps 4 >> $out 2>> gpurun_out/${tag}_scaling_n$N.err
fi
timeout 900 $TR --nproc-per-node $N tools/scaling.py --weak --width 16384 --height 8192 --iters 200 --halos 50,118 --steps 4 >> $out 2>> gpurun_out/${tag}_scaling_n$N.err
cat $out
timeout 600 $TR --nproc-per-node $N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; cut -c1-260 gpurun_out/${tag}_bench_n$N.json
