#!/usr/bin/env python
"""Phase timeline of ONE projection pass (GPU box): python tools/proj_timeline.py [rows T [W H]]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
cfg = baseline_config(1) if (W, H) == (1920, 1080) else baseline_config(1, width=W, height=H)
u, v, sm = synthetic_fields(W, H)
f = Fluid(cfg)
f.set_field("u", u); f.set_field("v", v)
f.set_option("projection_kernel", 1); f.set_option("autotune", 0)
f.set_option("temporal_block", T); f.set_option("tile_rows_per_warp", rows)
f.set_option("debug_timeline", 1)
for rep in range(3):
    f.stage_projection(T, 0.05)   # exactly one pass
    f.sync()
t = f.debug_timeline()
t0 = t[:, 0].min()
ent, ld, sw, st = [(t[:, k] - t0) / 1e3 for k in range(4)]
print(f"rows {rows} T {T}: {len(t)} tiles on {len(set(t[:,4]))} SMs; pass span {st.max():.1f} us")
print(f"  entry   : min {ent.min():6.1f} med {np.median(ent):6.1f} max {ent.max():6.1f} us")
print(f"  load    : med {np.median(ld-ent):6.1f} max {(ld-ent).max():6.1f} us   (entry -> tile in registers, tables built)")
print(f"  sweeps  : med {np.median(sw-ld):6.1f} min {(sw-ld).min():6.1f} max {(sw-ld).max():6.1f} us   ({2*T} half-sweeps; {np.median(sw-ld)/(2*T)*1e3:.0f} ns each)")
print(f"  store   : med {np.median(st-sw):6.1f} max {(st-sw).max():6.1f} us")
print(f"  CTA life: med {np.median(st-ent):6.1f} max {(st-ent).max():6.1f} us")
order = np.argsort(sw - ld)[::-1][:8]
tx = None
print("  slowest sweeps (tile index: us):", [(int(k), round(float((sw-ld)[k]), 1)) for k in order])
d = np.sort(sw - ld)[::-1]
print("  sweep time distribution (us), sorted:", [round(float(x), 1) for x in d[:24]], "... median", round(float(np.median(d)), 1))
