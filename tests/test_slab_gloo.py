"""Host-side logic of the y-slab driver (opensayal_b200/slab.py) on CPU: world_size-2 and -3 `gloo` process
groups exchanging ghost rows of a numpy stand-in whose operations have the same dependency reach as the real
stages (projection: one row per half-sweep; advection: a bounded gather).  The result must equal the
single-domain run exactly — it does only if the schedule exchanges often enough and addresses the right rows.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opensayal_b200 import slab as S

W, H, N_ITER, REACH = 24, 61, 7, 3


class NumpySlab:
    """Stand-in for FluidSlab: local rows [lo, hi) of a global H-row domain, ghost rows included."""

    def __init__(self, full, row0, rows, halo, first, last):
        self.row0, self.rows, self.halo, self.first, self.last = row0, rows, halo, first, last
        self.lo = row0 - (0 if first else halo)
        self.hi = row0 + rows + (0 if last else halo)
        self.f = {k: a[self.lo:self.hi].copy() for k, a in full.items()}  # ghosts start fresh
        self.width = W
        n = halo * W * 3
        self.send = [torch.empty(n, dtype=torch.float32) for _ in range(2)]
        self.recv = [torch.empty(n, dtype=torch.float32) for _ in range(2)]

    def _names(self, mask):
        return [n for bit, n in ((S.F_U, "u"), (S.F_V, "v"), (S.F_SMOKE, "smoke")) if mask & bit]

    def apply(self, op, source=None, d_t=None):
        kind = op[0]
        g = np.arange(self.lo, self.hi)[:, None]  # global row of every local row
        if kind == "forces":
            self.f["v"] += (0.01 * g).astype(np.float32)
        elif kind == "projection":
            own0 = self.row0 - self.lo
            for m in range(op[1]):
                # ("projection", k, depth): sweep only the rows that are still exact before iteration m of the chunk
                # (owned +- (depth - 2 m)); the others keep stale values, like rows outside the library's row window
                first, last = 0, len(self.f["u"])
                if len(op) > 2:
                    first = max(own0 - (op[2] - 2 * m), 0)
                    last = min(own0 + self.rows + (op[2] - 2 * m), last)
                for colour in (0, 1):
                    for name in ("u", "v"):
                        a = self.f[name]
                        new = a.copy()
                        for lr in range(first, last):
                            r = self.lo + lr
                            if r == 0 or r == H - 1 or (r + colour) % 2:
                                continue
                            up = a[lr - 1] if lr > 0 else 0.0          # missing neighbour => wrong, like a ghost edge
                            dn = a[lr + 1] if lr + 1 < len(a) else 0.0
                            new[lr] = np.float32(0.5) * a[lr] + np.float32(0.25) * (up + dn)
                        self.f[name] = new
        elif kind in ("advect_velocity", "advect_smoke"):
            names = ("u", "v") if kind == "advect_velocity" else ("smoke",)
            own0 = self.row0 - self.lo
            ghost = op[1] if len(op) > 1 else 0  # lazy schedule: also advect `ghost` rows beyond the owned ones
            first = max(own0 - ghost, 0)
            last = min(own0 + self.rows + ghost, len(self.f["u"]))
            for name in names:
                a = self.f[name]
                new = np.full_like(a, np.nan)  # rows that are not advected go stale, like the back buffer
                for lr in range(first, last):
                    r = self.lo + lr
                    for i in range(W):
                        src = min(max(r + (i % (2 * REACH + 1)) - REACH, 0), H - 1)
                        assert self.lo <= src < self.hi, "back-trace left the ghost rows"
                        new[lr, i] = a[src - self.lo, i]
                self.f[name] = new
        elif kind == "diffusion":  # in-place sweeps over u with one-row reach per colour, like apply_diffusion
            a = self.f["u"]
            for m in range(op[1]):
                for colour in (0, 1):
                    new = a.copy()
                    for lr in range(len(a)):
                        r = self.lo + lr
                        if r == 0 or r == H - 1 or (r + colour) % 2:
                            continue
                        up = a[lr - 1] if lr > 0 else 0.0
                        dn = a[lr + 1] if lr + 1 < len(a) else 0.0
                        new[lr] = (a[lr] + np.float32(0.125) * (up + dn)) / np.float32(1.25)
                    a = new
            self.f["u"] = a
        elif kind in ("extrapolation", "zero_pressure", "pressure_range"):
            pass
        else:
            raise ValueError(op)

    def pack(self, side, mask):
        own0 = self.row0 - self.lo
        rows = slice(own0, own0 + self.halo) if side == 0 else slice(own0 + self.rows - self.halo, own0 + self.rows)
        flat = np.concatenate([self.f[n][rows].ravel() for n in self._names(mask)])
        self.send[side][: flat.size] = torch.from_numpy(flat)
        return self.send[side][: flat.size]

    def recv_buffer(self, side, mask):
        return self.recv[side][: self.halo * W * len(self._names(mask))]

    def unpack(self, side, mask):
        own0 = self.row0 - self.lo
        rows = slice(own0 - self.halo, own0) if side == 0 else slice(own0 + self.rows, own0 + self.rows + self.halo)
        buf = self.recv[side].numpy()
        for k, n in enumerate(self._names(mask)):
            self.f[n][rows] = buf[k * self.halo * W:(k + 1) * self.halo * W].reshape(self.halo, W)

    def owned(self, name):
        own0 = self.row0 - self.lo
        return self.f[name][own0:own0 + self.rows]


def initial():
    rng = np.random.default_rng(3)
    return {k: rng.standard_normal((H, W)).astype(np.float32) for k in ("u", "v", "smoke")}


def single_domain(steps, viscous=False):
    s = NumpySlab(initial(), 0, H, 0, True, True)
    ops = [op[:1] + op[1:] if op[0] != "advect_velocity" else ("advect_velocity",)
           for op in S.step_schedule(N_ITER, 8, False, True, viscous=viscous) if op[0] != "exchange"]
    for _ in range(steps):
        for op in ops:
            s.apply(op)
    return s.f


def test_slab_rows_partition():
    for height, world in ((1080, 8), (61, 3), (16384, 8), (10, 4)):
        spans = [S.slab_rows(height, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == height
        for (a0, an), (b0, _) in zip(spans[:-1], spans[1:]):
            assert a0 + an == b0
        assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_schedule_shape():
    ops = S.step_schedule(50, 16, False, True)
    assert ops[0] == ("forces",)
    proj = [op for op in ops if op[0] == "projection"]
    assert sum(k for _, k in proj) == 50 and all(k <= 8 for _, k in proj)
    # every projection chunk is followed by an exchange of u and v
    for idx, op in enumerate(ops):
        if op[0] == "projection":
            assert ops[idx + 1] == ("exchange", S.F_U | S.F_V)
    assert ops[-1] == ("exchange", S.F_SMOKE) and ops[-2] == ("advect_smoke",)
    assert ("zero_pressure",) in S.step_schedule(4, 4, True, False)
    with pytest.raises(ValueError):
        S.step_schedule(4, 1, False, False)


@pytest.mark.parametrize("world,halo", [(2, 4), (3, 6), (4, 5)])
def test_local_exchange_matches_single_domain(world, halo):
    full = initial()
    slabs = []
    for r in range(world):
        row0, rows = S.slab_rows(H, world, r)
        slabs.append(NumpySlab(full, row0, rows, halo, r == 0, r == world - 1))
    ops = S.step_schedule(N_ITER, halo, False, True)
    for _ in range(3):
        S.run_schedule_local(slabs, ops)
    want = single_domain(3)
    for s in slabs:
        for name in ("u", "v", "smoke"):
            assert np.array_equal(s.owned(name), want[name][s.row0:s.row0 + s.rows]), (name, s.row0)


@pytest.mark.parametrize("world,halo", [(2, 4), (2, 6), (3, 9), (2, 18), (4, 5)])
def test_lazy_schedule_matches_single_domain(world, halo):
    """The schedule the library runs natively (exchanges only when the ghost depth runs out + one at the end of the
    step).  halo 18 >= 2*7 + 3 + 1: a single exchange per step."""
    full = initial()
    slabs = []
    for r in range(world):
        row0, rows = S.slab_rows(H, world, r)
        slabs.append(NumpySlab(full, row0, rows, halo, r == 0, r == world - 1))
    ops = S.lazy_schedule(N_ITER, halo, False, True, margin=REACH)
    if halo == 18:
        assert sum(op[0] == "exchange" for op in ops) == 1
    for _ in range(3):
        S.run_schedule_local(slabs, ops)
    want = single_domain(3)
    for s in slabs:
        for name in ("u", "v", "smoke"):
            assert np.array_equal(s.owned(name), want[name][s.row0:s.row0 + s.rows]), (name, s.row0)


@pytest.mark.parametrize("world,halo", [(2, 6), (3, 9), (2, 18), (3, 18), (4, 5)])
def test_shrinking_row_window_matches_single_domain(world, halo):
    """The library's linked slabs sweep, in iteration m of a projection chunk, only the owned rows +- (depth - 2 m):
    ghost rows beyond that are already wrong.  Same owned rows as the single domain, bit for bit."""
    full = initial()
    slabs = []
    for r in range(world):
        row0, rows = S.slab_rows(H, world, r)
        slabs.append(NumpySlab(full, row0, rows, halo, r == 0, r == world - 1))
    ops = S.lazy_schedule(N_ITER, halo, False, True, margin=REACH, windows=True)
    assert all(len(op) == 3 and op[2] >= 2 * op[1] for op in ops if op[0] == "projection")
    for _ in range(3):
        S.run_schedule_local(slabs, ops)
    want = single_domain(3)
    for s in slabs:
        for name in ("u", "v", "smoke"):
            assert np.array_equal(s.owned(name), want[name][s.row0:s.row0 + s.rows]), (name, s.row0)


@pytest.mark.parametrize("world,halo,T", [(2, 6, None), (3, 8, None), (2, 18, None), (3, 14, 7), (4, 5, 2), (2, 4, 1)])
def test_push_schedule_matches_single_domain(world, halo, T):
    """Push mode (the library's default for linked slabs): a pass of `it` iterations sweeps the owned rows +- 2 it and
    the edge tiles hand the new edge rows to the neighbours — as data, a windowed projection followed by an exchange."""
    full = initial()
    slabs = []
    for r in range(world):
        row0, rows = S.slab_rows(H, world, r)
        slabs.append(NumpySlab(full, row0, rows, halo, r == 0, r == world - 1))
    ops = S.push_schedule(N_ITER, halo, False, True, margin=REACH, temporal_block=T)
    proj = [op for op in ops if op[0] == "projection"]
    assert sum(op[1] for op in proj) == N_ITER and all(op[2] == 2 * op[1] <= halo for op in proj)
    for _ in range(3):
        S.run_schedule_local(slabs, ops)
    want = single_domain(3)
    for s in slabs:
        for name in ("u", "v", "smoke"):
            assert np.array_equal(s.owned(name), want[name][s.row0:s.row0 + s.rows]), (name, s.row0)


def test_push_schedule_shape():
    assert S.push_temporal_block(50, 18) == 8 and S.push_temporal_block(50, 16) == 8 and S.push_temporal_block(50, 10) == 5
    assert S.push_temporal_block(200, 18) == 8 and S.push_temporal_block(7, 18) == 7 and S.push_temporal_block(9, 18) == 5
    ops = S.push_schedule(50, 18, False, True)
    assert [op[1] for op in ops if op[0] == "projection"] == [8, 7, 7, 7, 7, 7, 7]
    assert ops[-1] == ("exchange", S.F_U | S.F_V | S.F_SMOKE)
    with pytest.raises(ValueError):
        S.push_schedule(50, 18, False, True, temporal_block=10)  # 20 ghost rows per pass > 18
    with pytest.raises(ValueError):
        S.push_schedule(50, 12, False, True, margin=16)


@pytest.mark.parametrize("schedule", ["step", "push"])
def test_viscous_schedules_run_the_diffusion_stage(schedule):
    """fluid.viscosity != 0 (the reference's shipped default): the slab schedules carry the diffusion sweeps with an
    exchange of u every halo // 2 sweeps, like step_impl does; same result as the single domain."""
    world, halo = 3, 6
    full = initial()
    slabs = []
    for r in range(world):
        row0, rows = S.slab_rows(H, world, r)
        slabs.append(NumpySlab(full, row0, rows, halo, r == 0, r == world - 1))
    ops = (S.step_schedule(N_ITER, halo, False, True, viscous=True) if schedule == "step"
           else S.push_schedule(N_ITER, halo, False, True, margin=REACH, viscous=True))
    assert sum(op[1] for op in ops if op[0] == "diffusion") == N_ITER
    for _ in range(2):
        S.run_schedule_local(slabs, ops)
    want = single_domain(2, viscous=True)
    plain = single_domain(2)
    assert not np.array_equal(want["u"], plain["u"]), "the stand-in's diffusion must change u"
    for s in slabs:
        for name in ("u", "v", "smoke"):
            assert np.array_equal(s.owned(name), want[name][s.row0:s.row0 + s.rows]), (name, s.row0)


def test_too_narrow_window_is_detected():
    """Sanity of the stand-in: a window that shrinks into rows the owned ones still depend on must change the result."""
    full = initial()
    slabs = []
    for r in range(2):
        row0, rows = S.slab_rows(H, 2, r)
        slabs.append(NumpySlab(full, row0, rows, 18, r == 0, r == 1))
    ops = [(op[0], op[1], op[2] - 2 * op[1] - 1) if op[0] == "projection" else op
           for op in S.lazy_schedule(N_ITER, 18, False, True, margin=REACH, windows=True)]
    S.run_schedule_local(slabs, ops)
    want = single_domain(1)
    assert any(not np.array_equal(s.owned(n), want[n][s.row0:s.row0 + s.rows]) for s in slabs for n in ("u", "v"))


def test_lazy_schedule_shape():
    ops = S.lazy_schedule(50, 120, False, True)
    assert [op[0] for op in ops] == ["forces", "projection", "extrapolation", "advect_velocity", "advect_smoke", "exchange"]
    assert ops[-1] == ("exchange", S.F_U | S.F_V | S.F_SMOKE) and ops[1] == ("projection", 50)
    ops = S.lazy_schedule(50, 32, True, False)
    proj = [op[1] for op in ops if op[0] == "projection"]
    assert proj == [16, 16, 16, 2] and sum(op[0] == "exchange" for op in ops) == 4  # 3 inside + 1 at the end
    assert ops[-1] == ("exchange", S.F_U | S.F_V)
    ops = S.lazy_schedule(7, 16, False, True, margin=3)  # depth left after 7 iterations: 2 < margin + 1
    assert [op[0] for op in ops].count("exchange") == 2
    with pytest.raises(ValueError):
        S.lazy_schedule(4, 8, False, False, margin=16)


def test_lazy_schedule_with_too_small_margin_is_detected():
    """The stand-in gathers from REACH rows away: a schedule built for margin 1 must trip its assertion."""
    full = initial()
    slabs = []
    for r in range(2):
        row0, rows = S.slab_rows(H, 2, r)
        slabs.append(NumpySlab(full, row0, rows, 8, r == 0, r == 1))
    ops = S.lazy_schedule(N_ITER, 8, False, True, margin=1)
    S.run_schedule_local(slabs, ops)
    want = single_domain(1)
    assert any(not np.array_equal(s.owned(n), want[n][s.row0:s.row0 + s.rows]) for s in slabs for n in ("u", "v"))


def test_too_thin_halo_is_detected_by_this_test_design():
    """Sanity of the stand-in: exchanging too rarely (halo claims 8, schedule thinks 16) must change results."""
    full = initial()
    slabs = []
    for r in range(2):
        row0, rows = S.slab_rows(H, 2, r)
        slabs.append(NumpySlab(full, row0, rows, 4, r == 0, r == 1))
    ops = S.step_schedule(N_ITER, 16, False, True)  # schedule for a deeper halo than the slabs hold
    S.run_schedule_local(slabs, ops)
    want = single_domain(1)
    assert any(not np.array_equal(s.owned("u"), want["u"][s.row0:s.row0 + s.rows]) for s in slabs)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, halo, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        row0, rows = S.slab_rows(H, world, rank)
        s = NumpySlab(initial(), row0, rows, halo, rank == 0, rank == world - 1)
        ops = S.step_schedule(N_ITER, halo, False, True)
        for _ in range(steps):
            S.run_schedule_dist(s, ops, rank, world)
        out[rank] = {n: s.owned(n).copy() for n in ("u", "v", "smoke")}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_exchange_matches_single_domain(world):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, 6, 2, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        want = single_domain(2)
        for r in range(world):
            row0, rows = S.slab_rows(H, world, r)
            for name in ("u", "v", "smoke"):
                assert np.array_equal(out[r][name], want[name][row0:row0 + rows]), (r, name)


def test_slab_schedule_choice_is_a_function_of_shared_numbers():
    """Thin slabs of a small grid recompute a deep halo (one exchange per step), everything else pushes; every rank
    derives the same answer from (width, rows per slab, n, world, margin)."""
    for world in (2, 4, 8):
        assert S.choose_slab_schedule(1920, 1080, 50, world) == ("deep", 118)      # bench.py --gpus N
        assert S.choose_slab_schedule(16384, 16384 // world, 50, world) == ("push", 18)  # BASELINE configs[3]
        assert S.choose_slab_schedule(16384, 8192, 200, world) == ("push", 18)     # BASELINE configs[4]
    assert S.choose_slab_schedule(1920, 400, 50, 4) == ("push", 18)                # too thin to hide the deep exchange
    assert S.choose_slab_schedule(1920, 1080, 50, 4, margin=6) == ("deep", 108)


def test_slab_rows_with_edge_bonus_partition_the_domain():
    for height, world, bonus in ((4320, 4, 69), (8640, 8, 69), (1000, 3, 7), (2160, 2, 69), (999, 4, 0)):
        parts = [S.slab_rows(height, world, r, bonus) for r in range(world)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == height
        assert all(parts[k][0] + parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        if bonus and world > 2:
            assert parts[0][1] - parts[1][1] in (bonus, bonus + 1, bonus - 1)
