#!/usr/bin/env python
"""Is the projection's speed data dependent?  Whole-domain 1920x1198, fixed plan, different initial fields (1 GPU)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

W, H, n = 1920, 1198, 50
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
u0, v0, sm = synthetic_fields(W, H)
cases = {"synthetic": (u0, v0), "zeros": (np.zeros_like(u0), np.zeros_like(v0)),
         "bottom 118 rows zero": (np.concatenate([u0[:1080], np.zeros_like(u0[1080:])]), np.concatenate([v0[:1080], np.zeros_like(v0[1080:])])),
         "tiny 1e-30": (u0 * 1e-30, v0 * 1e-30), "denormal 1e-40": ((u0.astype(np.float64) * 1e-40).astype(np.float32), (v0.astype(np.float64) * 1e-40).astype(np.float32))}
for plan in ("8x9", "12x4"):
    for label, (u, v) in cases.items():
        cfg = baseline_config(1, width=W, height=H)
        f = Fluid(cfg, device=0)
        r, t = plan.split("x")
        f.set_option("autotune", 0); f.set_option("temporal_block", int(t)); f.set_option("tile_rows_per_warp", int(r))
        st = torch.cuda.ExternalStream(f.stream)
        evs = []
        for k in range(12):
            f.set_field("u", u); f.set_field("v", v)   # the same input every time: projection converges otherwise
            with torch.cuda.stream(st):
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); f.stage_projection(n, 0.05); b.record(st); evs.append((a, b))
        f.sync()
        ms = sorted(x.elapsed_time(y) for x, y in evs[2:])
        print(f"plan {plan} {label:22s}: {ms[len(ms)//2]*1e3:7.1f} us", flush=True)
        f.close()
