"""tools/resident_timeline.py [rows T] — phase times of the resident projection's ring exchanges (GPU box).

Six globaltimer stamps per tile and sweep block (option debug_timeline): sweeps begin, sweeps done, ring stored, flag
published, neighbours' flags seen, halo loaded.  Prints the median / max over tiles of each phase, per block."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opensayal_b200 import Fluid  # noqa: E402
from opensayal_b200.synthetic import baseline_config, synthetic_fields  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = baseline_config(1)
c = cfg.c
u, v, sm = synthetic_fields(c.width, c.height)
f = Fluid(cfg)
for n, a in (("u", u), ("v", v), ("smoke", sm)):
    f.set_field(n, a)
f.set_option("resident", 2)
f.set_option("temporal_block", T)
f.set_option("tile_rows_per_warp", rows)
f.set_option("autotune", 0)
for _ in range(3):
    f.stage_projection(c.proj_n, c.d_t)
f.sync()
f.set_option("debug_timeline", 1)
f.stage_projection(c.proj_n, c.d_t)
f.sync()
raw = f.debug_timeline().reshape(-1)
blocks = -(-c.proj_n // T)
tiles = len(raw) // (blocks * 6)
t = raw[: tiles * blocks * 6].reshape(tiles, blocks, 6).astype(np.float64)
t0 = t[:, 0, 0].min()
print(f"rows {rows} T {T}: {tiles} tiles, {blocks} blocks; whole kernel (first stamp to last sweep end) "
      f"{(t[:, -1, 1].max() - t0) / 1e3:.1f} us; tile start spread {(t[:, 0, 0].max() - t0) / 1e3:.1f} us")
names = ["sweeps", "ring stores", "fence+publish", "wait", "halo loads"]
print("block   " + "  ".join(f"{n:>22s}" for n in names) + "     block period (med)")
for b in range(blocks - 1):
    d = [t[:, b, k + 1] - t[:, b, k] for k in range(5)]
    period = np.median(t[:, b + 1, 0] - t[:, b, 0])
    if b < 6 or b >= blocks - 3:
        print(f"{b:5d}   " + "  ".join(f"med {np.median(x) / 1e3:6.2f} max {x.max() / 1e3:6.2f}" for x in d) + f"   {period / 1e3:6.2f}")
d = [(t[:, :-1, k + 1] - t[:, :-1, k]) for k in range(5)]
print("all     " + "  ".join(f"med {np.median(x) / 1e3:6.2f} max {x.max() / 1e3:6.2f}" for x in d))
f.close()
