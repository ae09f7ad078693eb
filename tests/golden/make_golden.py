#!/usr/bin/env python
"""Generates the committed golden fixtures by EXECUTING THE REFERENCE'S OWN CODE.

Runs on a B200 box (`gpurun -- python tests/golden/make_golden.py gpurun_out/golden`): oracle/_ref/libsayal_ref.so
is /root/reference/src/fluid.cu + helper.cu compiled unmodified for sm_100 with the reference's Release flags
(oracle/Makefile).  For every case below the reference `Fluid` is constructed, loaded with the deterministic
synthetic fields, stepped with `Fluid::update`, and its masks and fields after each step are written to
`<case>.npz`.  The files are then copied into tests/golden/ and committed; the CPU suite checks the oracle
against them (tests/test_golden.py), the GPU suite checks the CUDA path.

The reference is --use_fast_math, so fields are compared within 1e-5 relative L2 per step (restarting from the
stored state each step); masks are compared bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from opensayal_b200.synthetic import baseline_config, synthetic_fields  # noqa: E402

STEPS = 3


def cases():
    """name -> Config.  Small grids (the fixtures are committed), every stage of Fluid::update exercised."""
    out = {}
    # BASELINE configs[0] scaled: closed gravity tank with pressure
    out["tank_64x36"] = baseline_config(0, width=64, height=36)
    # BASELINE configs[1] scaled: wind tunnel + disc + smoke, drain on
    c = baseline_config(1, width=96, height=54)
    c["sim.wind_tunnel.speed"] = 40.0
    out["tunnel_96x54"] = c
    # configs[2] flavour: smoke decay, several smoke stripes, pipe walls, pressure on, odd sizes (ragged rows)
    c = baseline_config(2, width=101, height=67)
    c["sim.projection.n"] = 20
    c["sim.wind_tunnel.speed"] = 30.0
    c["sim.wind_tunnel.pipe_height"] = 24
    c["sim.wind_tunnel.pipe_length"] = 30
    c["sim.wind_tunnel.smoke_length"] = 3
    c["sim.wind_tunnel.smoke_count"] = 3
    c["sim.wind_tunnel.smoke_height"] = 2
    c["sim.enable_pressure"] = 1
    c["sim.obstacle.center_x"] = 50
    c["sim.obstacle.center_y"] = 33
    c["sim.obstacle.radius"] = 6.5
    out["stripes_101x67"] = c
    return out


def main(outdir):
    from oracle.oracle import RefSim
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    for name, cfg in cases().items():
        c = cfg.c
        ref = RefSim(c, device=0)
        u, v, sm = synthetic_fields(c.width, c.height)
        ref.set_field("u", u)
        ref.set_field("v", v)
        ref.set_field("smoke", sm)
        if c.enable_pressure:
            ref.set_field("p", np.zeros_like(u))
        data = {"is_solid": ref.get_field("is_solid").astype(np.int8),
                "total_s": ref.get_field("total_s").astype(np.int8),
                "config_bytes": np.frombuffer(bytes(c), dtype=np.uint8)}
        names = ("u", "v", "smoke") + (("p",) if c.enable_pressure else ())
        for k in range(1, STEPS + 1):
            ref.step(None)
            for n in names:
                data[f"{n}_{k}"] = ref.get_field(n)
            if c.enable_pressure:
                data[f"prange_{k}"] = np.array(ref.pressure_range(), np.float32)
        np.savez_compressed(outdir / f"{name}.npz", **data)
        ref.close()
        print(name, {k: (a.shape, str(a.dtype)) for k, a in data.items() if k.endswith("_1") or k == "is_solid"})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
