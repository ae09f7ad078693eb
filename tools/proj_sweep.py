#!/usr/bin/env python
"""Time the projection stage for (kernel, rows-per-warp, T) plans on one grid (GPU box).
usage: python tools/proj_sweep.py [W H n] [--kernels 1,2] [--rows 8,10,12] [--T 4,5,6,7,8,10,12]"""
import sys
sys.path.insert(0, ".")
import torch
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

def arg(name, default):
    return [int(x) for x in (sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default).split(",")]

pdl = arg("--pdl", "1")[0]
pos = [a for a in sys.argv[1:] if a.isdigit() and sys.argv[sys.argv.index(a) - 1] not in ("--pdl", "--kernels", "--rows", "--T")]
W, H, n = (int(pos[0]), int(pos[1]), int(pos[2])) if len(pos) >= 3 else (1920, 1080, 50)
cfg = baseline_config(1, width=W, height=H)
if (W, H) == (1920, 1080):
    cfg = baseline_config(1)
u, v, sm = synthetic_fields(W, H)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kernel in arg("--kernels", "1"):
    for rows in arg("--rows", "8,10,12"):
        for T in arg("--T", "4,5,6,7,8,10,12"):
            f = Fluid(cfg)
            f.set_field("u", u); f.set_field("v", v)
            f.set_option("projection_kernel", kernel); f.set_option("autotune", 0)
            f.set_option("temporal_block", T); f.set_option("tile_rows_per_warp", rows)
            f.set_option("use_pdl", pdl)
            st = torch.cuda.ExternalStream(f.stream)
            try:
                for _ in range(3): f.stage_projection(n, 0.05)
            except Exception as e:
                print(f"kernel {kernel} rows {rows} T {T}: {e}"); f.close(); continue
            f.sync()
            l0 = f.launch_count
            reps = 10
            evs = []
            for _ in range(reps):
                with torch.cuda.stream(st): flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st); f.stage_projection(n, 0.05); b.record(st); evs.append((a, b))
            f.sync()
            ms = sorted(a.elapsed_time(b) for a, b in evs)
            passes = (f.launch_count - l0) // reps
            print(f"kernel {kernel} rows {rows:2d} T {T:2d}: {ms[len(ms)//2]*1e3:8.1f} us median, {ms[0]*1e3:8.1f} min, {passes} passes, {ms[len(ms)//2]*1e3/passes:6.1f} us/pass", flush=True)
            f.close()
