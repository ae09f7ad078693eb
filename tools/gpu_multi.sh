#!/bin/bash
# tools/gpu_multi.sh <tag> <N> — multi-GPU round on one box: IPC slab test on N real GPUs, slab probes, bench.
set -u
tag=${1:-r1}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 300 $TR tests/slab_ipc_worker.py > gpurun_out/${tag}_ipc_worker_n$N.log 2>&1; echo "ipc worker rc=$?"
tail -3 gpurun_out/${tag}_ipc_worker_n$N.log
timeout 600 $TR tools/slab_probe.py --halos 18,32,50,118 > gpurun_out/${tag}_probe_1080p_n$N.log 2>&1; grep "^N=" gpurun_out/${tag}_probe_1080p_n$N.log
timeout 900 $TR tools/slab_probe.py --width 16384 --rows-per-gpu $((16384 / N)) --halos 32,118 > gpurun_out/${tag}_probe_16k_n$N.log 2>&1; grep "^N=" gpurun_out/${tag}_probe_16k_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; cat gpurun_out/${tag}_bench_n$N.json
python bench.py --steps 50 --warmup 5 --skip-cpu-baseline > gpurun_out/${tag}_bench_n1_samebox.json 2>/dev/null; cut -c1-400 gpurun_out/${tag}_bench_n1_samebox.json
