"""y-slabs on the GPU: N slab sims (here: on one device, exchanging through device copies) must reproduce the
single-domain CUDA run bit for bit — projection is order-independent within a colour and advection is a gather
(SURVEY.md §8e).  The multi-process NCCL transport is the same schedule with send/recv instead of the copy."""
import numpy as np
import pytest

from opensayal_b200 import Fluid
from opensayal_b200 import slab as S
from opensayal_b200.synthetic import baseline_config, synthetic_fields

pytestmark = pytest.mark.gpu


def run_slabs(cfg, world, halo, steps, kernel=2):
    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height)
    slabs = []
    for r in range(world):
        row0, rows = S.slab_rows(c.height, world, r)
        s = S.FluidSlab(cfg, 0, row0, rows, halo, r == 0, r == world - 1)
        s.sim.set_option("projection_kernel", kernel)
        s.sim.set_field("u", u[row0:row0 + rows])
        s.sim.set_field("v", v[row0:row0 + rows])
        s.sim.set_field("smoke", sm[row0:row0 + rows])
        slabs.append(s)
    S.exchange_local(slabs, S.F_U | S.F_V | S.F_SMOKE)
    ops = S.step_schedule(c.proj_n, halo, bool(c.enable_pressure), bool(c.enable_smoke) and c.wt_smoke != 0)
    for _ in range(steps):
        S.run_schedule_local(slabs, ops)
    out = {n: np.concatenate([s.sim.get_field(n) for s in slabs]) for n in ("u", "v", "smoke", "p")}
    overflow = sum(s.sim.get_option("halo_overflow") for s in slabs)
    ranges = [(s.sim.min_pressure, s.sim.max_pressure) for s in slabs] if c.enable_pressure else None
    for s in slabs:
        s.close()
    return out, overflow, ranges


def run_single(cfg, steps, kernel=2):
    c = cfg.c
    f = Fluid(cfg)
    f.set_option("projection_kernel", kernel)
    u, v, sm = synthetic_fields(c.width, c.height)
    f.set_field("u", u)
    f.set_field("v", v)
    f.set_field("smoke", sm)
    for _ in range(steps):
        f.update(None)
    out = {n: f.get_field(n) for n in ("u", "v", "smoke", "p")}
    rng = (f.min_pressure, f.max_pressure) if c.enable_pressure else None
    f.close()
    return out, rng


@pytest.mark.parametrize("world,halo", [(2, 16), (3, 12), (4, 32)])
@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_slabs_bit_identical_to_single_gpu(world, halo, kernel):
    cfg = baseline_config(1, width=384, height=420)
    cfg["sim.projection.n"] = 20
    cfg["sim.wind_tunnel.speed"] = 60.0  # back-trace of 3 cells in x; rows stay well inside the halo
    want, _ = run_single(cfg, 3, kernel)
    got, overflow, _ = run_slabs(cfg, world, halo, 3, kernel)
    assert overflow == 0
    for n in ("u", "v", "smoke"):
        assert np.array_equal(got[n], want[n]), n


def test_slabs_with_pressure_and_gravity():
    cfg = baseline_config(0, width=256, height=288)
    cfg["sim.projection.n"] = 12
    want, wrange = run_single(cfg, 2)
    got, overflow, ranges = run_slabs(cfg, 3, 12, 2)
    assert overflow == 0
    for n in ("u", "v", "smoke", "p"):
        assert np.array_equal(got[n], want[n]), n
    assert min(r[0] for r in ranges) == wrange[0] and max(r[1] for r in ranges) == wrange[1]


def test_halo_overflow_is_reported():
    """A back-trace that leaves the ghost rows must be counted, not silently wrong."""
    cfg = baseline_config(1, width=256, height=256)
    cfg["sim.projection.n"] = 2
    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height, amplitude=400.0)  # 20-cell back-traces
    slabs = []
    for r in range(3):  # three slabs: the synthetic v has a node at H/2, not at H/3
        row0, rows = S.slab_rows(c.height, 3, r)
        s = S.FluidSlab(cfg, 0, row0, rows, 4, r == 0, r == 2)
        s.sim.set_field("u", u[row0:row0 + rows])
        s.sim.set_field("v", v[row0:row0 + rows])
        s.sim.set_field("smoke", sm[row0:row0 + rows])
        slabs.append(s)
    S.exchange_local(slabs, S.F_U | S.F_V | S.F_SMOKE)
    S.run_schedule_local(slabs, S.step_schedule(2, 4, False, True))
    assert sum(s.sim.get_option("halo_overflow") for s in slabs) > 0
