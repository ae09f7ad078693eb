#!/usr/bin/env python
"""BASELINE configs[0] (256x144 gravity tank, pressure on, n=50): us/step of the graph-replayed step, plain half-sweep
kernels vs the tiled kernel with pressure in shared memory (GPU box): python tools/config0_probe.py"""
import sys
sys.path.insert(0, ".")
import torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

for kern in (0, 1):
    cfg = baseline_config(0)
    f = Fluid(cfg)
    u, v, sm = synthetic_fields(cfg.c.width, cfg.c.height)
    f.set_field("u", u); f.set_field("v", v); f.set_field("smoke", sm)
    f.set_option("projection_kernel", kern)
    f.run(20); f.sync()
    st = torch.cuda.ExternalStream(f.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = f.launch_count
    e0.record(st); f.run(1000); e1.record(st); f.sync()
    ms = e0.elapsed_time(e1)
    print(f"configs[0] {cfg.c.width}x{cfg.c.height} pressure on, 1000 steps, projection_kernel={kern}: {ms:.1f} us/step, "
          f"{(f.launch_count - l0) // 1000} launches/step, {cfg.c.width * cfg.c.height / (ms * 1e-6) / 1e9:.3f} G cell-steps/s, "
          f"plan T={f.get_option('plan_temporal_block')} rows={f.get_option('plan_rows_per_warp')}", flush=True)
    f.close()
