#!/usr/bin/env python
"""A/B of an integer option on the bench workload, graph replay, L2-warm and L2-flushed (GPU box):
   python tools/fuse_ab.py [option=fuse_forces] [W H n]"""
import sys
sys.path.insert(0, ".")
import torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

opt = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "fuse_forces"
pos = [int(a) for a in sys.argv[1:] if a.isdigit()]
W, H, n = pos if len(pos) == 3 else (1920, 1080, 50)
cfg = baseline_config(1) if (W, H) == (1920, 1080) else baseline_config(1, width=W, height=H)
cfg["sim.projection.n"] = n
u, v, sm = synthetic_fields(W, H)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rnd in range(2):
    for val in (0, 1):
        f = Fluid(cfg)
        for name, a in (("u", u), ("v", v), ("smoke", sm)):
            f.set_field(name, a)
        f.set_option(opt, val)
        f.run(10); f.sync()
        st = torch.cuda.ExternalStream(f.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); f.run(200); e1.record(st); f.sync()
        warm = e0.elapsed_time(e1) / 200 * 1e3
        evs = []
        for _ in range(30):
            with torch.cuda.stream(st):
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); f.run(1); b.record(st); evs.append((a, b))
        f.sync()
        cold = sorted(x.elapsed_time(y) for x, y in evs)[15] * 1e3
        print(f"{opt}={val}: warm {warm:7.1f} us/step, flushed {cold:7.1f} us/step (median), plan T={f.get_option('plan_temporal_block')} rows={f.get_option('plan_rows_per_warp')}", flush=True)
        f.close()
