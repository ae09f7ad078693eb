"""Manual diagnostic (not a test): per-field one-step error of oracle and CUDA path vs the reference build."""
import sys
import numpy as np
sys.path.insert(0, ".")
from opensayal_b200 import Fluid
from opensayal_b200.synthetic import baseline_config, synthetic_fields
from oracle.oracle import OracleSim, RefSim


def rel(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b * b).sum()), 1e-30))


for label, cfg_fn, n in [("tunnel n=0", lambda: baseline_config(1, width=480, height=270), 0),
                         ("tunnel n=1", lambda: baseline_config(1, width=480, height=270), 1),
                         ("tunnel n=50", lambda: baseline_config(1, width=480, height=270), 50),
                         ("tank n=50", lambda: baseline_config(0), 50),
                         ("full n=50", lambda: baseline_config(1), 50)]:
    cfg = cfg_fn()
    cfg["sim.projection.n"] = n
    ref, cpu = RefSim(cfg.c), OracleSim(cfg.c)
    u, v, sm = synthetic_fields(cfg.c.width, cfg.c.height)
    for s in (ref, cpu):
        s.set_field("u", u); s.set_field("v", v); s.set_field("smoke", sm)
    print(label, "masks equal:", np.array_equal(ref.get_field("is_solid"), cpu.get_field("is_solid")),
          np.array_equal(ref.get_field("total_s"), cpu.get_field("total_s")))
    ref.step(None); cpu.step(None)
    for name in ("u", "v", "smoke") + (("p",) if cfg.c.enable_pressure else ()):
        r, c = ref.get_field(name), cpu.get_field(name)
        d = np.abs(r.astype(np.float64) - c)
        k = np.unravel_index(np.argmax(d), d.shape)
        print(f"   {name}: rel_l2 {rel(c, r):.3e}  max|d| {d.max():.3e} at row {k[0]} col {k[1]} (ref {r[k]:.6g} ours {c[k]:.6g})  n_bad(>1e-3) {(d > 1e-3).sum()}")
