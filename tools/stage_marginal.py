#!/usr/bin/env python
"""Marginal cost of each stage inside the replayed step graph (GPU box): the step is timed with one stage left
out at a time (option debug_skip; results are wrong, only the time is used).  python tools/stage_marginal.py [W H n]"""
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
for fuse in (0, 1):
    f = Fluid(cfg)
    for name, a in (("u", u), ("v", v), ("smoke", sm)):
        f.set_field(name, a)
    f.set_option("fuse_forces", fuse)
    f.run(10); f.sync()
    st = torch.cuda.ExternalStream(f.stream)
    res = {}
    for label, mask in (("full", 0), ("-forces", 1), ("-extrap", 2), ("-advect_v", 4), ("-advect_s", 8), ("-projection", 16),
                        ("only projection", 1 | 2 | 4 | 8), ("only advect", 1 | 2 | 16), ("full again", 0)):
        f.set_field("u", u); f.set_field("v", v); f.set_field("smoke", sm)
        f.set_option("debug_skip", mask)
        f.run(5); f.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); f.run(100); e1.record(st); f.sync()
        res[label] = e0.elapsed_time(e1) / 100 * 1e3
    print(f"fuse_forces={fuse}: " + ", ".join(f"{k} {v:.1f}" for k, v in res.items()), flush=True)
    f.close()
