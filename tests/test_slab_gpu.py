"""y-slabs on the GPU: N slab sims (here: on one device, exchanging through device copies) must reproduce the
single-domain CUDA run bit for bit — projection is order-independent within a colour and advection is a gather
(SURVEY.md §8e).  The multi-process NCCL transport is the same schedule with send/recv instead of the copy."""
import numpy as np
import pytest

from opensayal_b200 import Fluid
from opensayal_b200 import slab as S
from opensayal_b200.synthetic import baseline_config, synthetic_fields

pytestmark = pytest.mark.gpu


def run_slabs(cfg, world, halo, steps, kernel=1):
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


def run_single(cfg, steps, kernel=1):
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
@pytest.mark.parametrize("kernel", [0, 1])
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


# ---- native transport: peer-memory exchange + whole-step graph (csrc/slab_exchange.cu) ---------------------
def run_slabs_native(cfg, world, halo, steps, graph, margin=6, amplitude=40.0):
    """Slabs of one process on one device, linked with sayal_slab_connect_local; every slab runs the library's
    own slab schedule (sayal_step / sayal_run) and the exchange kernels talk through device memory."""
    from opensayal_b200 import Fluid as F
    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height, amplitude=amplitude)
    sims = []
    for r in range(world):
        row0, rows = S.slab_rows(c.height, world, r)
        f = F(cfg, device=0, slab=(row0, rows, halo))
        f.set_option("advect_margin", margin)  # back-traces: 60 * 0.05 = 3 cells (+ 2.25 synthetic) + stencil
        f.set_field("u", u[row0:row0 + rows])
        f.set_field("v", v[row0:row0 + rows])
        f.set_field("smoke", sm[row0:row0 + rows])
        sims.append(f)
    S.link_local(sims)
    for f in sims:
        # choose the tile plans now: the autotuner allocates / frees device memory, and cudaFree waits for the whole
        # device — with several slabs of ONE process on ONE device it would wait for a sibling's exchange kernel
        # that is itself waiting for this thread's next launch
        f.run(0)
    for f in sims:
        f.slab_exchange(S.F_U | S.F_V | S.F_SMOKE)
    if graph:
        for _ in range(steps):  # one replay per rank per step: every rank must be enqueued before any can finish
            for f in sims:
                f.run(1)
    else:
        for _ in range(steps):
            for f in sims:
                f.step_async(None)
    for f in sims:
        f.sync()
    out = {n: np.concatenate([f.get_field(n) for f in sims]) for n in ("u", "v", "smoke")}
    overflow = sum(f.get_option("halo_overflow") for f in sims)
    errors = sum(f.get_option("link_error") for f in sims)
    for f in sims:
        f.close()
    return out, overflow, errors


@pytest.mark.parametrize("world,halo", [(2, 16), (3, 12), (4, 32), (2, 47), (3, 60)])
@pytest.mark.parametrize("graph", [0, 1])
def test_native_linked_slabs_bit_identical_to_single_gpu(world, halo, graph):
    """halo 47 = 2 * 20 + 6 + 1 and above: a single exchange per step (hidden behind the smoke advection)."""
    cfg = baseline_config(1, width=384, height=420)
    cfg["sim.projection.n"] = 20
    cfg["sim.wind_tunnel.speed"] = 60.0
    want, _ = run_single(cfg, 3)
    got, overflow, errors = run_slabs_native(cfg, world, halo, 3, graph)
    assert errors == 0 and overflow == 0
    for n in ("u", "v", "smoke"):
        assert np.array_equal(got[n], want[n]), n


def test_native_linked_slabs_with_default_viscosity():
    """fluid.viscosity = 0.001 (the reference's shipped default, config_parser.cpp:117): the linked step carries the
    diffusion sweeps chunked by the halo with an exchange of u after each chunk; same bits as one GPU."""
    cfg = baseline_config(1, width=384, height=420)
    cfg["sim.projection.n"] = 20
    cfg["sim.wind_tunnel.speed"] = 60.0
    cfg["fluid.viscosity"] = 0.001
    want, _ = run_single(cfg, 2)
    got, overflow, errors = run_slabs_native(cfg, 3, 16, 2, graph=1)
    assert errors == 0 and overflow == 0
    for n in ("u", "v", "smoke"):
        assert np.array_equal(got[n], want[n]), n


@pytest.mark.parametrize("overlap", [0, 1])
def test_native_slabs_tall_domain_split_last_pass(overlap, monkeypatch):
    """Tall slabs (several tile rows each): with overlap on, the last projection pass of a chunk runs as edge tile
    rows + interior tile rows with the exchange in between on the aux stream; advection likewise.  Same bits."""
    monkeypatch.setenv("SAYAL_OVERLAP_EXCHANGE", str(overlap))
    cfg = baseline_config(1, width=256, height=1400)
    cfg["sim.projection.n"] = 21
    cfg["sim.wind_tunnel.speed"] = 60.0
    want, _ = run_single(cfg, 2)
    got, overflow, errors = run_slabs_native(cfg, 2, 20, 2, graph=1)
    assert errors == 0 and overflow == 0
    for n in ("u", "v", "smoke"):
        assert np.array_equal(got[n], want[n]), n


def test_native_slabs_report_a_too_small_margin():
    """Fast flow (20-cell back-traces) with advect_margin 6: gathers leave the rows known to be exact -> counted."""
    cfg = baseline_config(1, width=256, height=510)
    cfg["sim.projection.n"] = 4
    # three slabs: the synthetic v has a node at H/2, not at H/3
    _, overflow, errors = run_slabs_native(cfg, 3, 24, 1, graph=0, margin=6, amplitude=400.0)
    assert errors == 0 and overflow > 0


@pytest.mark.parametrize("world", [2, 3])
def test_linked_slabs_share_one_pressure_range_and_frame(world):
    """Fluid::min_pressure / max_pressure are ONE pair for the frame (fluid.cu:778-787, graphics_handler.cu:288-289):
    linked slabs reduce their ranges along the chain inside the step, so every slab reports the single-domain pair and
    the slabs' pressure frames, stacked, are the single-domain frame bit for bit."""
    from opensayal_b200 import Fluid as F
    cfg = baseline_config(0, width=256, height=288)  # gravity tank, enable_pressure
    cfg["sim.projection.n"] = 12
    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height)
    single = F(cfg, device=0)
    for n, a in (("u", u), ("v", v), ("smoke", sm)):
        single.set_field(n, a)
    single.run(3)
    single.sync()
    want_range = (single.min_pressure, single.max_pressure)
    want_frame = single.render_pixels()
    want_p = single.get_field("p")
    single.close()
    sims = []
    for r in range(world):
        row0, rows = S.slab_rows(c.height, world, r)
        f = F(cfg, device=0, slab=(row0, rows, 20))
        f.set_option("advect_margin", 6)
        for n, a in (("u", u), ("v", v), ("smoke", sm)):
            f.set_field(n, a[row0:row0 + rows])
        sims.append(f)
    S.link_local(sims)
    for f in sims:
        f.run(0)
    for f in sims:
        f.slab_exchange(S.F_U | S.F_V | S.F_SMOKE)
    for _ in range(3):
        for f in sims:
            f.run(1)
    for f in sims:
        f.sync()
    assert np.array_equal(np.concatenate([f.get_field("p") for f in sims]), want_p)
    for f in sims:
        assert (f.min_pressure, f.max_pressure) == want_range
    assert want_range[0] < 0.0 < want_range[1]
    frame = np.concatenate([f.render_pixels() for f in sims])
    assert np.array_equal(frame, want_frame)
    for f in sims:
        f.close()


def test_silent_neighbour_raises_a_sticky_link_error():
    """A linked slab whose neighbour never steps: the waits give up after 2 s, nothing is unpacked, and every later
    synchronising call reports SAYAL_ELINK instead of handing out fields (no silent corruption)."""
    from opensayal_b200 import SayalError
    from opensayal_b200._abi import SAYAL_ELINK
    cfg = baseline_config(1, width=256, height=256)
    cfg["sim.projection.n"] = 4
    sims = []
    for r in range(2):
        row0, rows = S.slab_rows(cfg.c.height, 2, r)
        sims.append(Fluid(cfg, device=0, slab=(row0, rows, 18)))
    S.link_local(sims)
    for f in sims:
        f.run(0)
    sims[0].step_async(None)  # sims[1] stays silent
    with pytest.raises(SayalError) as e:
        sims[0].sync()
    assert e.value.code == SAYAL_ELINK
    assert sims[0].get_option("link_error") != 0
    with pytest.raises(SayalError):
        sims[0].get_field("u")
    with pytest.raises(SayalError):
        sims[0].step_async(None)
    for f in sims:
        f.close()


def test_unlinked_slab_refuses_to_step():
    from opensayal_b200 import SayalError
    cfg = baseline_config(1, width=256, height=256)
    f = Fluid(cfg, slab=(64, 64, 8))
    with pytest.raises(SayalError):
        f.step_async(None)
    f.close()


def test_two_processes_ipc_link():
    """Two ranks (both on cuda:0, gloo rendez-vous) trade CUDA IPC handles and run the native slab schedule; rank 0
    compares the gathered rows with a single-domain run.  This is the multi-GPU path minus the second GPU."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    worker = Path(__file__).with_name("slab_ipc_worker.py")
    env = dict(os.environ, SAYAL_IPC_TEST_DEVICE="0", OMP_NUM_THREADS="1")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(worker)],
                          capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    assert "IPC-SLABS-OK" in proc.stdout
