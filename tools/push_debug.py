"""tools/push_debug.py — diagnose linked slabs of ONE process on ONE device (the layout of tests/test_slab_gpu.py).

    python tools/push_debug.py [--world 2] [--halo 16] [--steps 3] [--graph 0]

Runs the native slab schedule, then prints each sim's link_error, the control words of its link block and its private
counters (sayal_debug_link_words).  Options of the library are taken from the environment (SAYAL_USE_PDL,
SAYAL_SLAB_PUSH, ...), so the same scenario can be run under several settings from a shell loop.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opensayal_b200 import Fluid, SayalError  # noqa: E402
from opensayal_b200 import slab as S  # noqa: E402
from opensayal_b200.synthetic import baseline_config, synthetic_fields  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--halo", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--width", type=int, default=384)
    ap.add_argument("--height", type=int, default=420)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    cfg = baseline_config(1, width=a.width, height=a.height)
    cfg["sim.projection.n"] = a.iters
    cfg["sim.wind_tunnel.speed"] = 60.0
    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height)
    single = Fluid(cfg, device=0)
    for n, f in (("u", u), ("v", v), ("smoke", sm)):
        single.set_field(n, f)
    for _ in range(a.steps):
        single.step_async(None)
    single.sync()
    want = {n: single.get_field(n) for n in ("u", "v", "smoke")}
    single.close()

    sims = []
    for r in range(a.world):
        row0, rows = S.slab_rows(c.height, a.world, r)
        f = Fluid(cfg, device=0, slab=(row0, rows, a.halo))
        f.set_option("advect_margin", 6)
        for n, arr in (("u", u), ("v", v), ("smoke", sm)):
            f.set_field(n, arr[row0:row0 + rows])
        sims.append(f)
    S.link_local(sims)
    for f in sims:
        f.run(0)
    print("push_mode", [f.get_option("push_mode") for f in sims], "plan",
          [(f.get_option("plan_temporal_block"), f.get_option("plan_rows_per_warp")) for f in sims], flush=True)
    for f in sims:
        f.slab_exchange(S.F_U | S.F_V | S.F_SMOKE)
    for _ in range(a.steps):
        for f in sims:
            if a.graph:
                f.run(1)
            else:
                f.step_async(None)
    status = "ok"
    for k, f in enumerate(sims):
        try:
            f.sync()
        except SayalError as e:
            status = "LINK ERROR"
            print(f"sim {k}: {e}", flush=True)
    for k, f in enumerate(sims):
        w = f.debug_link_words()
        print(f"sim {k}: link_error {f.get_option('link_error')} xflag {w[0:2]} pflag {w[2:4]} piter {[hex(x) for x in w[4:6]]} "
              f"send_seq {w[64:66]} ticket {w[68:72]} range_seq {w[72]} step_seq {w[73]} push_ticket {w[74:76]}", flush=True)
    if status == "ok":
        got = {}
        try:
            got = {n: np.concatenate([f.get_field(n) for f in sims]) for n in ("u", "v", "smoke")}
        except SayalError as e:
            print("get_field:", e)
        for n in got:
            bad = int((got[n].view(np.uint32) != want[n].view(np.uint32)).sum())
            print(f"{n}: {bad} mismatched cells", flush=True)
            if bad:
                status = "MISMATCH"
    print("RESULT", status, flush=True)
    for f in sims:
        f.close()


if __name__ == "__main__":
    main()
