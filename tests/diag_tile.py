"""Manual diagnostic: tiled projection vs plain kernel for large temporal blocks / slab shapes."""
import sys
import numpy as np
sys.path.insert(0, ".")
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields

cfg = baseline_config(1, width=384, height=420)
u, v, sm = synthetic_fields(384, 420)


def run(kernel, iters, T=0, rows=0, slab=None, autotune=0):
    f = Fluid(cfg, slab=slab)
    f.set_option("projection_kernel", kernel)
    f.set_option("autotune", autotune)
    if T: f.set_option("temporal_block", T)
    if rows: f.set_option("tile_rows_per_warp", rows)
    r0, n = (0, 420) if slab is None else (slab[0], slab[1])
    f.set_field("u", u[r0:r0 + n]); f.set_field("v", v[r0:r0 + n])
    f.stage_projection(iters, 0.05)
    out = f.get_field("u"), f.get_field("v"), f.get_option("plan_temporal_block"), f.get_option("plan_rows_per_warp")
    f.close()
    return out

ref = run(0, 16)
for rows in (8, 10, 12):
    for T in (9, 12, 13, 14, 15, 16):
        try:
            a = run(1, 16, T, rows)
            bad = int((a[0] != ref[0]).sum() + (a[1] != ref[1]).sum())
            where = np.argwhere(a[0] != ref[0])[:3].tolist() if bad else []
            print(f"rows={rows} T={T}: bad={bad} plan=({a[2]},{a[3]}) {where}")
        except Exception as e:
            print(f"rows={rows} T={T}: {e}")
a = run(1, 16, autotune=1)
print("autotune:", int((a[0] != ref[0]).sum()), a[2], a[3])
