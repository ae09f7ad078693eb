#!/bin/bash
# tools/gpu_push_debug.sh — one-GPU diagnosis of linked slabs in one process (tools/push_debug.py) under several settings,
# then the GPU tests without -x.
mkdir -p gpurun_out
out=gpurun_out/push_debug.log
: > $out
run() { echo "=== $*" >> $out; env "$@" timeout 120 python tools/push_debug.py >> $out 2>&1; }
run A=default
run SAYAL_SLAB_PUSH=0
echo "=== graph=1" >> $out; timeout 120 python tools/push_debug.py --graph 1 >> $out 2>&1
echo "=== world=3 halo=12" >> $out; timeout 120 python tools/push_debug.py --world 3 --halo 12 >> $out 2>&1
echo "=== world=4 halo=16 1920x1080 n=50" >> $out; timeout 120 python tools/push_debug.py --world 4 --halo 18 --width 1920 --height 1080 --iters 50 >> $out 2>&1
cat $out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/push_debug_pytest.log 2>&1; tail -15 gpurun_out/push_debug_pytest.log
