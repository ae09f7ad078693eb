#!/usr/bin/env python
"""Per-stage CUDA-event timing of one step on the bench workload (GPU box): python tools/step_breakdown.py [W H n]"""
import sys
sys.path.insert(0, ".")
import torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

pos = [int(a) for a in sys.argv[1:] if a.isdigit()]
W, H, n = pos if len(pos) == 3 else (1920, 1080, 50)
cfg = baseline_config(1) if (W, H) == (1920, 1080) else baseline_config(1, width=W, height=H)
cfg["sim.projection.n"] = n
u, v, sm = synthetic_fields(W, H)
for adv in (0, 1, 2):
    f = Fluid(cfg)
    for name, a in (("u", u), ("v", v), ("smoke", sm)):
        f.set_field(name, a)
    f.set_option("advect_kernel", adv)
    f.run(5); f.sync()
    st = torch.cuda.ExternalStream(f.stream)
    d_t = cfg.c.d_t
    stages = [("forces", lambda: f.stage_forces(None, d_t)), ("projection", lambda: f.stage_projection(n, d_t)),
              ("extrapolation", f.stage_extrapolation), ("advect_velocity", lambda: f.stage_advect_velocity(d_t)),
              ("advect_smoke", lambda: f.stage_advect_smoke(d_t))]
    acc = {k: [] for k, _ in stages}
    for rep in range(12):
        for k, fn in stages:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); fn(); b.record(st); acc[k].append((a, b))
    f.sync()
    print(f"advect_kernel={adv}: " + ", ".join(f"{k} {sorted(x.elapsed_time(y) for x, y in v)[len(v)//2]*1e3:.1f} us" for k, v in acc.items()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); f.run(50); e1.record(st); f.sync()
    print(f"   graph replay: {e0.elapsed_time(e1)/50*1e3:.1f} us/step, plan T={f.get_option('plan_temporal_block')} rows={f.get_option('plan_rows_per_warp')}")
    f.close()
