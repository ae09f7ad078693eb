#!/bin/bash
# tools/gpu_weak_config4.sh <N> — BASELINE configs[4]: weak scaling, 16384 x 8192 cells per GPU, n = 200 (pushing passes),
# and configs[2] (3840 x 2160, n = 100, decay) on one GPU when N = 1.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1500 $TR tools/scaling.py --weak --width 16384 --height 8192 --iters 200 --halos 18 --steps 4 2> gpurun_out/weak4_n$N.err | tee gpurun_out/weak4_n$N.jsonl
if [ "$N" = "1" ]; then
python - <<'PY' | tee gpurun_out/config2_n1.json
import json, sys
sys.path.insert(0, ".")
import torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields
cfg = baseline_config(2)
c = cfg.c
f = Fluid(cfg)
u, v, sm = synthetic_fields(c.width, c.height)
for n, a in (("u", u), ("v", v), ("smoke", sm)):
    f.set_field(n, a)
st = torch.cuda.ExternalStream(f.stream)
f.run(5); f.sync()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st); f.run(20); b.record(st); f.sync()
ms = a.elapsed_time(b) / 20
print(json.dumps({"config": "BASELINE configs[2]: 3840x2160, n=100, smoke decay 0.05", "ms_per_step": ms,
                  "cell_steps_per_s": c.width * c.height / (ms * 1e-3), "algorithmic_bytes_per_cell_step": 8 + 100 * 17 + 34,
                  "achieved_GBps_algorithmic": c.width * c.height * 1742 / (ms * 1e-3) / 1e9,
                  "plan": [f.get_option("plan_temporal_block"), f.get_option("plan_rows_per_warp")]}))
PY
fi
