#!/usr/bin/env python
"""Projection stage on an UNLINKED slab sim vs a whole-domain sim with the same number of local rows (1 GPU):
   python tools/slab_proj_probe.py [W rows halo n]"""
import sys
sys.path.insert(0, ".")
import torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

pos = [int(a) for a in sys.argv[1:] if a.isdigit()]
W, rows, halo, n = pos if len(pos) == 4 else (1920, 1080, 118, 50)
H = 2 * rows
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(f, label):
    st = torch.cuda.ExternalStream(f.stream)
    for _ in range(3):
        f.stage_projection(n, 0.05)
    f.sync()
    evs = []
    for _ in range(10):
        with torch.cuda.stream(st):
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); f.stage_projection(n, 0.05); b.record(st); evs.append((a, b))
    f.sync()
    ms = sorted(x.elapsed_time(y) for x, y in evs)
    print(f"{label}: {ms[5]*1e3:7.1f} us  plan T={f.get_option('plan_temporal_block')} rows={f.get_option('plan_rows_per_warp')} local_rows={f.get_option('local_rows')}", flush=True)

for mode in ("auto", "8x9", "12x4", "10x10"):
    cfg = baseline_config(1, width=W, height=H)
    cfg["sim.wind_tunnel.pipe_height"] = H // 4
    u, v, sm = synthetic_fields(W, H, rows=(0, rows))
    f = Fluid(cfg, device=0, slab=(0, rows, halo))
    f.set_field("u", u); f.set_field("v", v)
    if mode != "auto":
        r, t = mode.split("x")
        f.set_option("autotune", 0); f.set_option("temporal_block", int(t)); f.set_option("tile_rows_per_warp", int(r))
    timed(f, f"slab  {W}x{rows}+{halo} {mode:6s}")
    f.close()
    cfg = baseline_config(1, width=W, height=rows + halo)
    u, v, sm = synthetic_fields(W, rows + halo)
    f = Fluid(cfg, device=0)
    f.set_field("u", u); f.set_field("v", v)
    if mode != "auto":
        r, t = mode.split("x")
        f.set_option("autotune", 0); f.set_option("temporal_block", int(t)); f.set_option("tile_rows_per_warp", int(r))
    timed(f, f"whole {W}x{rows + halo} {mode:6s}")
    f.close()
