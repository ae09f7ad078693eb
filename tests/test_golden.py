"""Golden fixtures produced by EXECUTING THE REFERENCE (tests/golden/make_golden.py ran oracle/_ref — the
reference's own fluid.cu, unmodified, sm_100, Release flags — on a B200; the .npz files are its masks and its
fields after 1, 2, 3 calls of Fluid::update from the deterministic synthetic start).

CPU suite: the oracle restatement against them (this is what pins the oracle without a GPU).
GPU suite: the CUDA path against them, through the C ABI.

Bar: masks bit for bit; u, v, p, smoke within 1e-5 relative L2 per step (the reference is --use_fast_math, both
of ours are IEEE), every step restarted from the stored reference state; the cells whose value depends on the
reference's racy extrapolation faces (H4) are excluded with the same window as tests/test_reference_parity.py.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from opensayal_b200 import Fluid
from opensayal_b200._abi import SayalConfig
from opensayal_b200.synthetic import synthetic_fields
from oracle.oracle import OracleSim

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLDEN.glob("*.npz"))
TOL = 1e-5
RADIUS = 12


def rel_l2(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / den) if den > 0 else float(np.abs(a).max())


def exclude_contested(arrays, H, W):
    r = lambda j: H - 1 - j
    for (row, col) in [(r(0), 1), (r(H - 1), 1), (r(1), 0), (r(1), W - 1)]:
        r0, r1 = max(row - RADIUS, 0), min(row + RADIUS + 1, H)
        c0, c1 = max(col - RADIUS, 0), min(col + RADIUS + 1, W)
        for a in arrays:
            a[r0:r1, c0:c1] = 0


def load_case(name):
    d = np.load(GOLDEN / f"{name}.npz")
    cfg = SayalConfig.from_buffer_copy(d["config_bytes"].tobytes())
    return d, cfg


def check_against_golden(make_sim, step, name):
    d, cfg = load_case(name)
    H, W = cfg.height, cfg.width
    sim = make_sim(cfg)
    assert np.array_equal(sim.get_field("is_solid"), d["is_solid"].astype(np.int32)), "is_solid"
    assert np.array_equal(sim.get_field("total_s"), d["total_s"].astype(np.int32)), "total_s"
    names = ("u", "v", "smoke") + (("p",) if cfg.enable_pressure else ())
    u, v, sm = synthetic_fields(W, H)
    state = {"u": u, "v": v, "smoke": sm, "p": np.zeros_like(u)}
    worst = 0.0
    for k in (1, 2, 3):
        for n in names:
            sim.set_field(n, state[n])
        step(sim)
        for n in names:
            want, got = d[f"{n}_{k}"].copy(), sim.get_field(n)
            state[n] = d[f"{n}_{k}"]
            exclude_contested((want, got), H, W)
            e = rel_l2(got, want)
            worst = max(worst, e)
            assert e <= TOL, f"{name}: {n} after step {k}: rel L2 {e:.3e}"
    return worst


def test_fixtures_present():
    assert CASES == ["stripes_101x67", "tank_64x36", "tunnel_96x54"]
    for name in CASES:
        d, cfg = load_case(name)
        assert d["is_solid"].shape == (cfg.height, cfg.width)
        assert C.sizeof(SayalConfig) == d["config_bytes"].size


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    worst = check_against_golden(lambda cfg: OracleSim(cfg), lambda s: s.step(None), name)
    print(f"{name}: worst one-step rel L2 of the oracle vs the reference's output: {worst:.2e}")


@pytest.mark.parametrize("name", CASES)
def test_oracle_pressure_range_matches_golden(name):
    d, cfg = load_case(name)
    if not cfg.enable_pressure:
        pytest.skip("pressure disabled in this case")
    o = OracleSim(cfg)
    u, v, sm = synthetic_fields(cfg.width, cfg.height)
    for n, a in (("u", u), ("v", v), ("smoke", sm)):
        o.set_field(n, a)
    o.step(None)
    mn, mx = o.pressure_range()
    rmn, rmx = d["prange_1"]
    assert abs(mn - rmn) <= 1e-4 * abs(rmn) and abs(mx - rmx) <= 1e-4 * abs(rmx)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_golden(name):
    def make(cfg):
        from opensayal_b200 import Config
        return Fluid(Config(cfg))
    worst = check_against_golden(make, lambda s: s.update(None), name)
    print(f"{name}: worst one-step rel L2 of the CUDA path vs the reference's output: {worst:.2e}")
