"""Multi-GPU y-slabs for the step path.  No reference equivalent: OpenSayal is single-GPU.

Memory rows are y (index = (H-1-j)*W + i, fluid.cu:163-165), so a slab is one contiguous row range.  Rank k of
N owns rows [k*H/N, (k+1)*H/N) and keeps `halo` ghost rows of every field on each interior side.

Why ghost rows are enough (and keep the result bit-identical to one GPU):
  * projection (fluid.cu:229-262): a half-sweep has dependency radius one row, so after k iterations the ghost
    rows are wrong only within 2k rows of the local array edge.  With halo >= 2k the owned rows stay exact for
    k iterations; then neighbours swap edge rows.  `iters_per_exchange = halo // 2`.  The update is
    order-independent within a colour, so the owned rows hold the same bits as the single-GPU sweep.
  * advection (fluid.cu:560-617) is a gather: exact as long as the back-trace stays inside the ghost rows.  The
    kernels count every sample that would leave them (`halo_overflow`); a non-zero count is an error.
  * forces / extrapolation are functions of position and of the local neighbour row: applied to every held row.

Three schedules, all bit-identical to one GPU:
  * `step_schedule` (driven from Python, NCCL / gloo / device-copy exchanges): forces, [projection chunk,
    exchange(u,v)] x ceil(n / (halo/2)), extrapolation, velocity advection, exchange(u,v), smoke advection,
    exchange(smoke).
  * `lazy_schedule` (what the library runs natively for linked slabs, csrc/sayal_api.cu + csrc/slab_exchange.cu:
    peer-memory stores over NVLink from its own kernels, the whole step one CUDA-graph replay): exchanges only when
    the remaining ghost depth is too small for the next operation, plus one exchange of u, v, smoke at the end of
    the step — ONE exchange per step when halo >= 2 n + margin + 1, hidden behind the interior smoke advection.

  * `push_schedule` (the library's default for linked slabs, csrc/projection_pack.cu): every projection pass of
    `it` iterations sweeps the owned rows +- 2 it, writes the owned rows, and its edge tiles store the new edge rows
    straight into the neighbours' ghost rows — as data: [projection(it, depth 2 it), exchange(u, v)] per pass.  Ghost
    rows cost 2 T + margin instead of 2 n + margin; compute and exchange are one kernel.

The schedule is data (a list of ops), so the same program drives real slabs over torch.distributed, several
slabs on one GPU (tests) and a numpy stand-in under gloo on CPU (tests of the host logic).
"""
from __future__ import annotations

import os

from typing import Callable, List, Optional, Sequence, Tuple

F_U, F_V, F_SMOKE, F_P = 1, 2, 4, 8


def slab_rows(height: int, world: int, rank: int, edge_bonus: int = 0) -> Tuple[int, int]:
    """(row0, rows) of rank's slab: contiguous memory rows, remainder spread over the first ranks.
    edge_bonus > 0 (deep-halo schedule, world > 2): the two edge slabs own that many rows more than the interior ones —
    they sweep ghost rows on one side only, the interior slabs on two, so equal ownership leaves the edge GPUs waiting
    at the end-of-step exchange."""
    if edge_bonus > 0 and world > 2:
        interior = (height - 2 * edge_bonus) // world
        edge = interior + edge_bonus
        sizes = [edge] + [interior] * (world - 2) + [edge]
        for k in range(height - sum(sizes)):  # remainder: one row each, from the first rank on
            sizes[k % world] += 1
        return sum(sizes[:rank]), sizes[rank]
    base, extra = divmod(height, world)
    rows = base + (1 if rank < extra else 0)
    row0 = rank * base + min(rank, extra)
    return row0, rows


def _diffusion_ops(n_iterations: int, halo: int) -> List[tuple]:
    """apply_diffusion (fluid.cu:775-777): n sweeps over u when fluid.viscosity != 0.  A sweep reaches one row per
    colour like a projection half-sweep, so it is chunked the same way: halo // 2 sweeps, then an exchange of u —
    what step_impl (csrc/sayal_api.cu) does for linked slabs."""
    ops: List[tuple] = []
    per, done = halo // 2, 0
    while done < n_iterations:
        k = min(per, n_iterations - done)
        ops.append(("diffusion", k))
        ops.append(("exchange", F_U))
        done += k
    return ops


def step_schedule(n_iterations: int, halo: int, pressure: bool, smoke: bool, viscous: bool = False) -> List[tuple]:
    """One Fluid::update (fluid.cu:770-795) as slab operations.  viscous: fluid.viscosity != 0 (the reference's
    shipped default is 0.001, config_parser.cpp:117)."""
    if halo < 2:
        raise ValueError("halo must be >= 2 rows (one SOR iteration reaches two rows)")
    per = halo // 2
    ops: List[tuple] = [("forces",)]
    if pressure:
        ops.append(("zero_pressure",))
    if viscous:
        ops += _diffusion_ops(n_iterations, halo)
    done = 0
    while done < n_iterations:
        k = min(per, n_iterations - done)
        ops.append(("projection", k))
        ops.append(("exchange", F_U | F_V))
        done += k
    if pressure:
        ops.append(("pressure_range",))
    ops.append(("extrapolation",))
    ops.append(("advect_velocity",))
    ops.append(("exchange", F_U | F_V))
    if smoke:
        ops.append(("advect_smoke",))
        ops.append(("exchange", F_SMOKE))
    return ops


def lazy_schedule(n_iterations: int, halo: int, pressure: bool, smoke: bool, margin: int = 16,
                  windows: bool = False) -> List[tuple]:
    """The schedule the library runs for linked slabs (csrc/sayal_api.cu step_impl), as data.

    Ghost rows are exact to depth D beyond the owned rows.  An SOR iteration costs two rows of depth; an exchange
    restores D = halo.  Exchanges happen only when the next operation needs more depth than is left:
      * before an iteration when D < 2,
      * before the velocity advection when D < margin + 1 (it runs on the owned rows and ONE ghost row each side
        — the smoke sampler reads the new velocity one row away — and gathers from up to `margin` rows away),
    and once at the end of the step for u, v and smoke together.  With halo >= 2 n + margin + 1 that is the only
    exchange of the step.

    windows=True: projection ops carry the depth that is exact when the chunk starts, ("projection", k, depth) —
    iteration m of the chunk then only needs to sweep the owned rows +- (depth - 2 m), the library's shrinking row
    window (projection_pack.cu): rows beyond it are already wrong and nobody reads them before the next exchange."""
    if halo < margin + 1 or halo < 2:
        raise ValueError("halo must be >= margin + 1")
    ops: List[tuple] = [("forces",)]
    if pressure:
        ops.append(("zero_pressure",))
    depth, done = halo, 0
    while done < n_iterations:
        if depth < 2:
            ops.append(("exchange", F_U | F_V))
            depth = halo
        k = min(n_iterations - done, depth // 2)
        ops.append(("projection", k, depth) if windows else ("projection", k))
        depth -= 2 * k
        done += k
    if pressure:
        ops.append(("pressure_range",))
    ops.append(("extrapolation",))
    if depth < margin + 1:
        ops.append(("exchange", F_U | F_V))
        depth = halo
    ops.append(("advect_velocity", 1))  # owned rows +- 1 ghost row
    if smoke:
        ops.append(("advect_smoke",))
    ops.append(("exchange", F_U | F_V | (F_SMOKE if smoke else 0)))
    return ops


def push_temporal_block(n_iterations: int, halo: int) -> int:
    """The iterations per pass every rank of a chain uses in push mode (tiled_push_temporal_block, projection_pack.cu):
    the even split of n into passes of at most min(halo // 2, 8)."""
    cap = min(halo // 2, 8)
    if cap < 1 or n_iterations <= 0:
        return 0
    passes = -(-n_iterations // cap)
    return -(-n_iterations // passes)


def push_schedule(n_iterations: int, halo: int, pressure: bool, smoke: bool, margin: int = 16,
                  temporal_block: Optional[int] = None, viscous: bool = False) -> List[tuple]:
    """The library's default schedule for linked slabs (push mode), as data.  Pass k of `it` iterations sweeps the
    owned rows +- 2 it ghost rows (("projection", it, 2 it): the shrinking window of that pass) and its edge tiles
    deliver the new edge rows to the neighbours, which the stand-in expresses as an exchange of u and v after the
    pass.  Needs halo >= max(2 T, margin + 1)."""
    T = temporal_block or push_temporal_block(n_iterations, halo)
    if halo < margin + 1 or (n_iterations > 0 and (T < 1 or 2 * T > halo)):
        raise ValueError("push mode needs halo >= max(2 T, margin + 1)")
    ops: List[tuple] = [("forces",)]
    if pressure:
        ops.append(("zero_pressure",))
    if viscous:
        ops += _diffusion_ops(n_iterations, halo)
    if n_iterations > 0:
        passes = -(-n_iterations // T)
        base, longer = divmod(n_iterations, passes)
        for k in range(passes):
            it = base + (1 if k < longer else 0)
            ops.append(("projection", it, 2 * it))
            ops.append(("exchange", F_U | F_V))
    if pressure:
        ops.append(("pressure_range",))
    ops.append(("extrapolation",))
    ops.append(("advect_velocity", 1))
    if smoke:
        ops.append(("advect_smoke",))
    ops.append(("exchange", F_U | F_V | (F_SMOKE if smoke else 0)))
    return ops


def n_fields(mask: int) -> int:
    return bin(mask & 0xF).count("1")


class FluidSlab:
    """Adapter: one opensayal_b200.Fluid slab + its torch pack buffers."""

    def __init__(self, cfg, device: int, row0: int, rows: int, halo: int, first: bool, last: bool):
        import torch

        from .fluid import Fluid
        self.torch = torch
        self.sim = Fluid(cfg, device=device, slab=(row0, rows, halo))
        self.device = torch.device("cuda", device)
        self.halo, self.width, self.rows, self.row0 = halo, cfg.c.width, rows, row0
        self.first, self.last = first, last  # no neighbour on the low-row / high-row side
        n = halo * cfg.c.width * 3
        with torch.cuda.device(self.device):
            self.send = [torch.empty(n, dtype=torch.float32, device=self.device) for _ in range(2)]
            self.recv = [torch.empty(n, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.stream = torch.cuda.ExternalStream(self.sim.stream, device=self.device)
        self.d_t = cfg.c.d_t

    # ---- stages -------------------------------------------------------------------------------
    def apply(self, op: tuple, source=None, d_t: Optional[float] = None):
        d_t = self.d_t if d_t is None else d_t
        s = self.sim
        kind = op[0]
        if kind == "forces":
            s.stage_forces(source, d_t)
        elif kind == "zero_pressure":
            s.stage_zero_pressure()
        elif kind == "projection":
            s.stage_projection(op[1], d_t)
        elif kind == "diffusion":
            s.stage_diffusion(op[1], d_t)
        elif kind == "pressure_range":
            pass  # stage_projection already queued the local range; the global one is an all-reduce (pressure_range())
        elif kind == "extrapolation":
            s.stage_extrapolation()
        elif kind == "advect_velocity":
            s.stage_advect_velocity(d_t)
        elif kind == "advect_smoke":
            s.stage_advect_smoke(d_t)
        else:
            raise ValueError(op)

    def pack(self, side: int, mask: int):
        self.sim.pack_edge(side, self.halo, mask, self.send[side].data_ptr())
        return self.send[side][: self.halo * self.width * n_fields(mask)]

    def recv_buffer(self, side: int, mask: int):
        return self.recv[side][: self.halo * self.width * n_fields(mask)]

    def unpack(self, side: int, mask: int):
        self.sim.unpack_ghost(side, self.halo, mask, self.recv[side].data_ptr())

    def close(self):
        self.sim.close()


def exchange_local(slabs: Sequence, mask: int) -> None:
    """Neighbour exchange between slabs living in ONE process (tests; single-GPU emulation)."""
    for a, b in zip(slabs[:-1], slabs[1:]):  # a is above b: a's high-row edge <-> b's low-row ghost
        src = a.pack(1, mask)
        dst = b.recv_buffer(0, mask)
        _copy(a, b, dst, src)
        b.unpack(0, mask)
        src = b.pack(0, mask)
        dst = a.recv_buffer(1, mask)
        _copy(b, a, dst, src)
        a.unpack(1, mask)


def _copy(src_slab, dst_slab, dst, src):
    torch = getattr(src_slab, "torch", None)
    if torch is not None and src.is_cuda:
        # order: packed on src stream -> copy -> unpack on dst stream
        ev = torch.cuda.Event()
        ev.record(src_slab.stream)
        dst_slab.stream.wait_event(ev)
        with torch.cuda.stream(dst_slab.stream):
            dst.copy_(src, non_blocking=True)
    else:
        dst[...] = src


def exchange_dist(slab, mask: int, rank: int, world: int, group=None) -> None:
    """Neighbour exchange over torch.distributed (NCCL between GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    ops = []
    torch = getattr(slab, "torch", None)
    ctx = torch.cuda.stream(slab.stream) if torch is not None and getattr(slab, "stream", None) is not None else _null()
    with ctx:
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, slab.pack(0, mask), rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, slab.recv_buffer(0, mask), rank - 1, group))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, slab.pack(1, mask), rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, slab.recv_buffer(1, mask), rank + 1, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if rank > 0:
            slab.unpack(0, mask)
        if rank < world - 1:
            slab.unpack(1, mask)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run_schedule_local(slabs: Sequence, ops: Sequence[tuple], source=None, d_t=None) -> None:
    for op in ops:
        if op[0] == "exchange":
            exchange_local(slabs, op[1])
        else:
            for s in slabs:
                s.apply(op, source, d_t)


def run_schedule_dist(slab, ops: Sequence[tuple], rank: int, world: int, source=None, d_t=None, group=None) -> None:
    for op in ops:
        if op[0] == "exchange":
            exchange_dist(slab, op[1], rank, world, group)
        else:
            slab.apply(op, source, d_t)


def link_local(slabs: Sequence) -> None:
    """Connect slabs that live in ONE process (same or peer-accessible devices) for the native exchange."""
    sims = [getattr(s, "sim", s) for s in slabs]
    for a, b in zip(sims[:-1], sims[1:]):  # a holds the rows above b
        a.connect_local(1, b)
        b.connect_local(0, a)


def link_dist(sim, rank: int, world: int, group=None) -> None:
    """Trade CUDA IPC handles of the neighbour-writable blocks with the two neighbour ranks (once)."""
    import torch.distributed as dist
    mine = sim.ipc_export()
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    if rank > 0:
        sim.ipc_connect(0, everyone[rank - 1])
    if rank < world - 1:
        sim.ipc_connect(1, everyone[rank + 1])
    dist.barrier(group=group)


def choose_slab_schedule(width: int, rows_per_slab: int, n_iterations: int, world: int, margin: int = 16):
    """("push" | "deep", halo) for linked slabs of `rows_per_slab` owned rows — a function of numbers every rank knows,
    so all ranks choose alike.  Deep halo (2 n + margin + 2 ghost rows, recomputed, one exchange per step) costs the
    projection time x the share of ghost rows swept (they shrink by two rows per iteration: half the halo on average,
    plus a tile row's worth of quantisation); push mode (thin halo, every pass stores its edge rows into the
    neighbour) costs a fixed ~35 us per step on a B200 (measured, profiles/README.md: system fences of the pushing
    tiles, pass flags, the wait kernel).  Thin slabs of a small grid take the deep halo, everything else pushes."""
    deep_halo = 2 * n_iterations + margin + 2
    push_halo = max(margin + 2, 16)
    if rows_per_slab < 4 * deep_halo:  # the deep schedule hides its exchange under interior rows: needs room
        return "push", push_halo
    projection_us = rows_per_slab * width * n_iterations * 1.3e-6
    sides = 2 if world > 2 else 1
    deep_extra = projection_us * sides * (deep_halo / 2 + 10) / rows_per_slab
    return ("deep", deep_halo) if deep_extra < 35.0 else ("push", push_halo)


class SlabFluid:
    """`Fluid` over N y-slabs, one process per slab (torch.distributed must be initialised).

    transport "p2p" (default): ghost rows travel as peer-memory stores over NVLink issued by the library's own
    kernels on the sim's stream (csrc/slab_exchange.cu) and the whole step is one CUDA-graph replay per rank.
    transport "nccl": the same schedule driven from here, edge rows packed and sent with NCCL send/recv."""

    def __init__(self, cfg, rank: int, world: int, device: int, halo: Optional[int] = None, group=None,
                 transport: str = "p2p", margin: int = 16, push: Optional[bool] = None, balance: bool = False):
        self.cfg, self.rank, self.world, self.group = cfg, rank, world, group
        self.transport = transport if world > 1 else "none"
        c = cfg.c
        if halo is None:  # thin halo + pushing passes, or deep halo + one exchange per step (choose_slab_schedule)
            mode, halo = choose_slab_schedule(c.width, c.height // max(world, 1), c.proj_n, world, margin)
            if push is None:
                push = mode == "push"
        self.push = push  # None: the library's default (push whenever the halo allows)
        # deep halo: an interior slab sweeps about halo / 2 + 10 ghost rows more than an edge slab (see
        # choose_slab_schedule); `balance` lets the edge slabs own that many rows more
        self.edge_bonus = halo // 2 + 10 if (balance and push is False and world > 2) else 0
        row0, rows = slab_rows(c.height, world, rank, self.edge_bonus)
        if world > 1 and rows < halo:
            raise ValueError(f"slab of {rows} rows is thinner than the halo ({halo})")
        self.row0, self.rows, self.halo = row0, rows, halo
        self.slab = FluidSlab(cfg, device, row0, rows, halo if world > 1 else 0, rank == 0, rank == world - 1)
        self.ops = step_schedule(c.proj_n, halo, bool(c.enable_pressure), bool(c.enable_smoke) and c.wt_smoke != 0,
                                 viscous=c.viscosity != 0) if world > 1 else None
        if self.transport == "p2p":
            self.sim.set_option("advect_margin", margin)  # rows the advection may gather from (see lazy_schedule)
            if push is not None and not os.environ.get("SAYAL_SLAB_PUSH"):
                self.sim.set_option("slab_push", 1 if push else 0)
            link_dist(self.sim, rank, world, group)
            self.sim.run(0)  # choose the projection tile plans (may time candidates) before the first exchange

    @property
    def sim(self):
        return self.slab.sim

    def set_initial(self, u, v, smoke) -> None:
        """Owned rows of each field (rows x W); ghosts are filled by an exchange."""
        self.sim.set_field("u", u)
        self.sim.set_field("v", v)
        self.sim.set_field("smoke", smoke)
        if self.transport == "p2p":
            self.sim.slab_exchange(F_U | F_V | F_SMOKE)
        elif self.world > 1:
            exchange_dist(self.slab, F_U | F_V | F_SMOKE, self.rank, self.world, self.group)

    def update(self, source=None, d_t=None) -> None:
        if self.transport in ("none", "p2p"):
            self.sim.step_async(source, d_t)
        else:
            run_schedule_dist(self.slab, self.ops, self.rank, self.world, source, d_t, self.group)

    def run(self, steps: int, d_t=None) -> None:
        """`steps` updates with an inactive source; one graph replay per step on the native transports."""
        if self.transport in ("none", "p2p"):
            self.sim.run(steps, d_t)
        else:
            for _ in range(steps):
                self.update(None, d_t)

    def sync(self) -> None:
        self.sim.sync()

    def halo_overflow(self) -> int:
        return self.sim.get_option("halo_overflow")

    def pressure_range(self):
        """Fluid::min_pressure / max_pressure of the whole domain.  Linked slabs get it from the library (reduced along
        the chain inside the step); the NCCL transport all-reduces the per-slab pairs here."""
        mn, mx = self.sim.min_pressure, self.sim.max_pressure
        if self.world == 1 or self.transport == "p2p":
            return mn, mx
        import torch
        import torch.distributed as dist
        t = torch.tensor([mn, -mx], dtype=torch.float32, device=self.slab.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return float(t[0]), float(-t[1])

    def close(self):
        self.slab.close()


# ----------------------------------------------------------------------------------------------------
GATED_CHUNK = 40  # steps enqueued behind one gate: stays far below the driver's launch-queue depth


def gated_steps(sim, steps, one_step, events, barrier):
    """Run `steps` timed steps so that no host work can fall inside the timed region: per chunk, hold the stream with
    the host-released gate, enqueue two untimed alignment steps and the chunk's timed steps, meet the other ranks at a
    HOST barrier (every rank has enqueued everything), then open the gate.  The device then runs the whole chunk from
    its queue; a host thread that stalls afterwards cannot make a neighbour wait for rows."""
    done = 0
    while done < steps:
        n = min(GATED_CHUNK, steps - done)
        sim.stream_hold()
        one_step()
        one_step()
        for k in range(done, done + n):
            one_step(*events[k])
        barrier()
        sim.stream_release()
        sim.sync()
        done += n


def device_fields(width, height, rows, device):
    """Synthetic u, v, smoke of memory rows [r0, r0 + n) generated on the GPU (torch), for grids where the numpy
    generator of synthetic.py would take longer than the measurement (16384 x 16384).  Same modes, no noise term."""
    import math

    import torch
    r0, n = rows
    r = torch.arange(r0, r0 + n, dtype=torch.float64, device=device)[:, None]
    y = (height - 1 - r) + 0.5
    x = torch.arange(width, dtype=torch.float64, device=device)[None, :] + 0.5
    two_pi = 2.0 * math.pi
    cx, sx = torch.cos(two_pi * 3 * x / width), torch.sin(two_pi * 3 * x / width)
    u = (40.0 * sx * torch.cos(two_pi * 2 * y / height)).float()
    v = (-40.0 * cx * torch.sin(two_pi * 2 * y / height)).float()
    smoke = (0.5 * (1.0 + torch.sin(two_pi * 8 * x / width) * torch.sin(two_pi * 8 * y / height))).float()
    return u.contiguous(), v.contiguous(), smoke.contiguous()


def strong_16384_record(world, rank, local, steps=8, width=16384, height=16384, iters=50, with_single=True):
    """BASELINE configs[3]: 16384 x 16384, n = 50, strong scaling over `world` y-slabs (north_star: >= 85 % parallel
    efficiency at 8 GPUs).  Returns, on rank 0, {"ms_per_step": ..., "n1_ms_per_step": ...}: the time of the N-slab
    run and — measured in the same process on rank 0's GPU after the slabs are gone — of the single-GPU run."""
    import torch
    import torch.distributed as dist

    from .fluid import Fluid
    from .synthetic import baseline_config

    def config():
        cfg = baseline_config(3, width=width, height=height)
        cfg["sim.projection.n"] = iters
        return cfg

    def timed(sim_like, sim, barrier):
        stream = torch.cuda.ExternalStream(sim.stream)
        sim_like.run(3)
        sim_like.sync()
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]

        def one_step(a=None, b=None):
            if a is not None:
                a.record(stream)
            sim_like.run(1)
            if b is not None:
                b.record(stream)

        gated_steps(sim, steps, one_step, events, barrier)
        return sum(a.elapsed_time(b) for a, b in events) / steps

    def load(sim, row0, rows):
        u, v, sm = device_fields(width, height, (row0, rows), torch.device("cuda", local))
        torch.cuda.current_stream().synchronize()
        for name, t in (("u", u), ("v", v), ("smoke", sm)):
            sim.set_field_device(name, t.data_ptr())
        sim.sync()  # the tensors may go once the copies are done

    out = {"grid": [width, height], "sor_iterations": iters, "steps": steps, "scaling": "strong", "n_gpus": world}
    if world > 1:
        sf = SlabFluid(config(), rank, world, local)
        load(sf.sim, sf.row0, sf.rows)
        sf.sim.slab_exchange(F_U | F_V | F_SMOKE)
        ms = timed(sf, sf.sim, dist.barrier)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bad = torch.tensor([sf.halo_overflow(), sf.sim.get_option("link_error")], dtype=torch.int64, device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        out.update(ms_per_step=float(t[0]), halo_rows=sf.halo, halo_overflow=int(bad[0]), link_error=int(bad[1]),
                   push_mode=bool(sf.sim.get_option("push_mode")),
                   plan=[sf.sim.get_option("plan_temporal_block"), sf.sim.get_option("plan_rows_per_warp")])
        sf.close()
        dist.barrier()
    if rank == 0 and (with_single or world == 1):
        f = Fluid(config(), device=local)
        load(f, 0, height)
        ms1 = timed(f, f, lambda: None)
        out["n1_ms_per_step"] = ms1
        out["n1_plan"] = [f.get_option("plan_temporal_block"), f.get_option("plan_rows_per_warp")]
        if world == 1:
            out["ms_per_step"] = ms1
        f.close()
    if world > 1:
        dist.barrier()
    if "ms_per_step" in out:
        out["cell_steps_per_s"] = width * height / (out["ms_per_step"] * 1e-3)
    if world > 1 and "n1_ms_per_step" in out:
        out["speedup_vs_1gpu_same_box"] = out["n1_ms_per_step"] / out["ms_per_step"]
    return out if rank == 0 else None


def slab_parity_vs_single(cfg, sf, rank, world, local, steps=3):
    """The multi-GPU test proper, inside the bench (the GPU test box has one GPU): restart the slabs from the synthetic
    fields, run `steps` updates, gather the owned rows on rank 0 and compare them BIT FOR BIT with the same domain
    stepped on one GPU.  Returns a dict on rank 0."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from .fluid import Fluid
    from .synthetic import synthetic_fields

    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height, rows=(sf.row0, sf.rows))
    sf.set_initial(u, v, sm)
    sf.run(steps)
    sf.sync()
    names = ("u", "v", "smoke")
    mine = {n: sf.sim.get_field(n) for n in names}
    status = (sf.halo_overflow(), sf.sim.get_option("link_error"))
    parts = [None] * world
    dist.gather_object((mine, status), parts if rank == 0 else None, dst=0)
    result = None
    if rank == 0:
        f = Fluid(cfg, device=local)
        gu, gv, gs = synthetic_fields(c.width, c.height)
        for n, a in (("u", gu), ("v", gv), ("smoke", gs)):
            f.set_field(n, a)
        f.run(steps)
        f.sync()
        mismatched = {}
        for n in names:
            got = np.concatenate([p[0][n] for p in parts])
            want = f.get_field(n)
            # bit-for-bit (NaN-safe): compare the words, not the values
            mismatched[n] = int((got.view(np.uint32) != want.view(np.uint32)).sum())
        f.close()
        result = {"steps": steps, "fields": list(names), "mismatched_cells": mismatched,
                  "bit_identical": all(m == 0 for m in mismatched.values()),
                  "halo_overflow": int(sum(p[1][0] for p in parts)), "link_error": int(max(p[1][1] for p in parts))}
    dist.barrier()
    return result


def bench_slabs(args, workload_config: Callable, config_block: Callable, ClockSampler, measured_hbm_peak: Callable,
                algorithmic_bytes_per_cell_step: Callable):
    """bench.py's N > 1 leg: weak scaling, one rank per GPU, launched by torch.distributed.run."""
    import gc
    import os
    import time

    import torch
    import torch.distributed as dist

    from .synthetic import synthetic_fields

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        # NCCL prints its version banner on stdout: keep stdout for the one JSON line
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    cfg = workload_config(world)
    c = cfg.c
    margin = int(os.environ.get("SAYAL_ADVECT_MARGIN", "16"))
    # thin halo + pushing passes or deep halo + one exchange per step: choose_slab_schedule (1080-row slabs of this grid
    # take the deep halo; the 16384^2 strong-scaling record below pushes).  SAYAL_SLAB_HALO / SAYAL_SLAB_PUSH override.
    halo = int(os.environ.get("SAYAL_SLAB_HALO", "0")) or None
    transport = os.environ.get("SAYAL_SLAB_TRANSPORT", "p2p")
    sf = SlabFluid(cfg, rank, world, local, halo=halo, transport=transport, margin=margin,
                   balance=bool(os.environ.get("SAYAL_SLAB_BALANCE")))  # measured: no effect at 1080 rows per slab
    halo = sf.halo
    u, v, sm = synthetic_fields(c.width, c.height, rows=(sf.row0, sf.rows))
    sf.set_initial(u, v, sm)
    stream = sf.slab.stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    warm = max(args.warmup, 3)

    no_flush = bool(os.environ.get("SAYAL_BENCH_NO_FLUSH"))  # experiments only: the L2-warm step, same harness

    def one_step(a=None, b=None):
        if not no_flush:
            with torch.cuda.stream(stream):
                flush.fill_(1)
        if a is not None:
            a.record(stream)
        sf.run(1)
        if b is not None:
            b.record(stream)

    # warm-up runs the very loop body that is timed below: the first call of anything (torch's fill kernel, an
    # event) loads code lazily and can stall a rank for milliseconds — which its neighbours would then book as a
    # slow first step while they wait for its rows
    for _ in range(warm):
        one_step(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    sf.sync()
    dist.barrier()
    torch.cuda.synchronize()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    gc.collect()
    gc.disable()  # no collector pause in the enqueue loop
    launches0 = sf.sim.launch_count
    with ClockSampler(local, period=0.004) as clocks:
        gated_steps(sf.sim, args.steps, one_step, events, dist.barrier)
        torch.cuda.synchronize()
        dist.barrier()
    gc.enable()
    untimed = 2 * -(-args.steps // GATED_CHUNK)  # alignment steps
    launches = (sf.sim.launch_count - launches0) * args.steps // (args.steps + untimed)
    step_ms = [s.elapsed_time(e) for s, e in events]
    ms_local = sum(step_ms) / args.steps
    mine = torch.tensor(step_ms, dtype=torch.float64, device="cuda")
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    every = torch.stack(every).cpu()          # [rank][step]
    worst = every.max(dim=0).values           # slowest rank of every step
    per_step = sorted(float(x) for x in worst)
    k_bad = int(worst.argmax())
    slowest_step = {"index": k_bad, "ms": round(float(worst[k_bad]), 4),
                    "ms_per_rank": [round(float(x), 3) for x in every[:, k_bad]]}
    if os.environ.get("SAYAL_BENCH_DEBUG"):
        import sys
        print(f"[rank {rank}] step ms: min {min(step_ms):.4f} med {sorted(step_ms)[len(step_ms) // 2]:.4f} max {max(step_ms):.4f}; "
              f"first 8: {[round(x, 3) for x in step_ms[:8]]}", file=sys.stderr, flush=True)
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    bad = torch.tensor([sf.halo_overflow(), sf.sim.get_option("link_error")], dtype=torch.int64, device="cuda")
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    cells = c.width * c.height
    value = cells / (ms * 1e-3)

    # end to end: owned rows from pinned host buffers (one batched upload), K steps, owned rows back (one batched
    # download) — wall clock, max over ranks
    pinned = {k: torch.from_numpy(a).pin_memory() for k, a in (("u", u), ("v", v), ("smoke", sm))}
    outs = {k: torch.empty_like(t_).pin_memory() for k, t_ in pinned.items()}
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sf.sim.set_fields_from({k: t_.data_ptr() for k, t_ in pinned.items()})
    if transport == "p2p":
        sf.sim.slab_exchange(F_U | F_V | F_SMOKE)
    else:
        exchange_dist(sf.slab, F_U | F_V | F_SMOKE, rank, world, None)
    for _ in range(args.steps):
        sf.update()
    sf.sim.get_fields_into({k: t_.data_ptr() for k, t_ in outs.items()})
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = cells * args.steps / float(te[0])
    field_bytes = 3 * cells * 4
    launches_t = torch.tensor([launches], dtype=torch.int64, device="cuda")
    dist.all_reduce(launches_t, op=dist.ReduceOp.SUM)
    plan_t = torch.tensor([sf.sim.get_option("plan_temporal_block"), sf.sim.get_option("plan_rows_per_warp"),
                           sf.sim.get_option("local_rows"), sf.sim.get_option("push_mode")], dtype=torch.int64, device="cuda")
    plans = [torch.zeros_like(plan_t) for _ in range(world)]
    dist.all_gather(plans, plan_t)

    # N slabs == one GPU, bit for bit (rank 0 steps the stacked domain on its own GPU and compares)
    parity = None if os.environ.get("SAYAL_BENCH_SKIP_PARITY") else slab_parity_vs_single(cfg, sf, rank, world, local)
    sf.close()
    del flush
    torch.cuda.empty_cache()
    dist.barrier()
    strong = None
    if not os.environ.get("SAYAL_BENCH_SKIP_STRONG"):
        strong = strong_16384_record(world, rank, local)

    peak, peak_src = measured_hbm_peak()
    b_alg = algorithmic_bytes_per_cell_step(c.proj_n, int(bool(c.enable_pressure)), 1)
    line = None
    if rank == 0:
        line = {
            "metric": "cell-steps/sec (n=50 SOR)", "value": value, "unit": "cell-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(cfg, world),
            "slabs": {"halo_rows": halo, "push_mode": bool(int(plans[0][3])), "edge_slab_bonus_rows": sf.edge_bonus,
                      "exchange": (("projection passes store their edge rows into the neighbour's ghost rows over NVLink (one kernel "
                                    "computes and exchanges); " if bool(int(plans[0][3])) else
                                    "deep halo: ghost rows recomputed through the projection, no exchange inside it; ") +
                                   "end-of-step exchange kernel (peer-memory stores + flags) under the smoke advection; graph-replayed step")
                      if transport == "p2p" else "NCCL send/recv of packed edge rows",
                      "halo_overflow": int(bad[0]), "link_error": int(bad[1]),
                      "tile_plans_per_rank": [{"temporal_block": int(p[0]), "rows_per_warp": int(p[1]),
                                               "local_rows": int(p[2])} for p in plans]},
            "timing": "all ranks enqueue 2 alignment + K timed steps behind a host-released gate, meet at a host barrier, "
                      "then open the gate; CUDA events per step on every rank's stream, mean over steps, max over ranks",
            "roofline": {"bound": "hbm", "kernel": "whole step, all GPUs", "achieved": round(value * b_alg / 1e9, 1),
                         "peak": peak * world, "unit": "GB/s", "frac": round(value * b_alg / 1e9 / (peak * world), 4),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs"},
            "cpu_baseline": None,
            "e2e": {"value": e2e_value, "unit": "cell-steps/s", "h2d_bytes_per_step": field_bytes / args.steps + 20,
                    "d2h_bytes_per_step": field_bytes / args.steps},
            "gpu_launches": int(launches_t[0]), "clocks": clocks.summary(),
            "step_ms_slowest_rank": {"min": round(per_step[0], 4), "median": round(per_step[len(per_step) // 2], 4),
                                     "p90": round(per_step[int(0.9 * (len(per_step) - 1))], 4), "max": round(per_step[-1], 4),
                                     "slowest_step": slowest_step},
            "parity_vs_single_gpu": parity, "link_error": int(bad[1]),
            "strong_16384": strong,
        }
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and ((parity is not None and not parity["bit_identical"]) or int(bad[1]) or int(bad[0])):
        import json
        import sys
        print(json.dumps(line), flush=True)
        sys.exit(3)  # a throughput number over wrong fields is not a result
    return line
