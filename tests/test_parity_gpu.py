"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact.  Both sides execute the same IEEE-754 operation sequence, so integer flags
AND fp32 fields must be equal element for element (np.array_equal; -0.0 == +0.0 is accepted).
"""
import numpy as np
import pytest

from opensayal_b200 import Config, Fluid, Source
from opensayal_b200._abi import SayalSource
from opensayal_b200.synthetic import baseline_config, synthetic_fields
from oracle.oracle import OracleSim

pytestmark = pytest.mark.gpu


def pair(cfg, seed=1234, amplitude=40.0, fields=("u", "v", "smoke")):
    W, H = cfg.c.width, cfg.c.height
    gpu, cpu = Fluid(cfg), OracleSim(cfg.c)
    u, v, sm = synthetic_fields(W, H, seed=seed, amplitude=amplitude)
    for name, a in (("u", u), ("v", v), ("smoke", sm)):
        if name in fields:
            gpu.set_field(name, a)
            cpu.set_field(name, a)
    return gpu, cpu


def assert_same(gpu, cpu, names=("u", "v", "smoke"), what=""):
    for n in names:
        a, b = gpu.get_field(n), cpu.get_field(n)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            r, i = bad[0]
            raise AssertionError(f"{what} {n}: {len(bad)} cells differ, first at row {r} col {i}: "
                                 f"gpu {a[r, i]!r} oracle {b[r, i]!r}; max |diff| {np.nanmax(np.abs(a - b))}")


CONFIGS = {
    "tank_256x144": lambda: baseline_config(0),
    "tunnel_384x216": lambda: baseline_config(1, width=384, height=216),
    "odd_203x157": lambda: Config.defaults(203, 157, **{"fluid.viscosity": 0.0, "sim.wind_tunnel.speed": 60.0,
                                                          "sim.obstacle.radius": 17.5, "sim.physics.g": -3.0}),
    "closed_disc_130x170": lambda: Config.defaults(130, 170, **{"fluid.viscosity": 0.0, "sim.enable_drain": 0,
                                                                 "sim.obstacle.center_x": 40,
                                                                 "sim.obstacle.center_y": 120,
                                                                 "sim.obstacle.radius": 23.0}),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_masks_bit_exact(name):
    cfg = CONFIGS[name]()
    gpu, cpu = Fluid(cfg), OracleSim(cfg.c)
    assert np.array_equal(gpu.is_solid, cpu.get_field("is_solid"))
    assert np.array_equal(gpu.total_s, cpu.get_field("total_s"))


def test_masks_pipe_walls_and_full_size():
    cfg = baseline_config(1)
    cfg.c.wt_pipe_length = 300  # dead in the shipped parser (config_parser.cpp:64) but part of the formula
    gpu, cpu = Fluid(cfg), OracleSim(cfg.c)
    assert np.array_equal(gpu.is_solid, cpu.get_field("is_solid"))
    assert np.array_equal(gpu.total_s, cpu.get_field("total_s"))
    assert gpu.is_solid.sum() > 8967


def test_forces_bit_exact():
    cfg = baseline_config(1, width=384, height=216)
    cfg["sim.physics.g"] = -5.0
    cfg["fluid.drag_coeff"] = 0.3
    cfg["sim.wind_tunnel.smoke_count"] = 3
    cfg["sim.wind_tunnel.smoke_height"] = 10
    gpu, cpu = pair(cfg)
    src = Source(True, 0.7, 1.5, (200, 100))
    gpu.stage_forces(src, 0.05)
    cpu.forces(src._c(), 0.05)
    assert_same(gpu, cpu, what="forces")
    gpu.stage_forces(None, 0.02)
    cpu.forces(None, 0.02)
    assert_same(gpu, cpu, what="forces(inactive)")


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("kernel", [0, 1])
def test_projection_bit_exact(name, kernel):
    cfg = CONFIGS[name]()
    cfg["sim.enable_pressure"] = 0
    gpu, cpu = pair(cfg)
    gpu.set_option("projection_kernel", kernel)
    gpu.stage_projection(7, 0.05)
    cpu.projection(7, 0.05)
    assert_same(gpu, cpu, names=("u", "v"), what=f"projection kernel={kernel}")


@pytest.mark.parametrize("kernel", [1])
@pytest.mark.parametrize("rows", [8, 10, 12])
@pytest.mark.parametrize("T", [1, 2, 3, 4, 5, 8, 12, 16])
def test_tiled_projection_any_temporal_block(T, rows, kernel):
    """Tiling / temporal blocking must not change a bit: several tiles in x and y, n not divisible by T,
    every tile variant (rows per warp), both register-tile kernels."""
    cfg = baseline_config(1, width=520, height=470)
    gpu, cpu = pair(cfg)
    gpu.set_option("projection_kernel", kernel)
    gpu.set_option("temporal_block", T)
    gpu.set_option("tile_rows_per_warp", rows)
    gpu.stage_projection(13, 0.05)
    cpu.projection(13, 0.05)
    assert_same(gpu, cpu, names=("u", "v"), what=f"tiled T={T}")


@pytest.mark.parametrize("rows", [8, 10, 12])
@pytest.mark.parametrize("T", [1, 2, 3, 5, 8])
def test_resident_projection_any_block(T, rows):
    """Resident plans (the whole projection in one cooperative launch, tiles kept in registers, ring exchange between
    sweep blocks of T iterations) must give the bits of the plain half-sweeps: several tiles in x and y, n not divisible
    by T (uneven blocks), every tile variant, both exchange parities (odd and even numbers of blocks)."""
    cfg = baseline_config(1, width=520, height=470)
    for n in (13, 14):
        gpu, cpu = pair(cfg)
        gpu.set_option("resident", 2)  # resident plans only
        gpu.set_option("temporal_block", T)
        gpu.set_option("tile_rows_per_warp", rows)
        gpu.stage_projection(n, 0.05)
        assert gpu.get_option("plan_resident") == 1
        gpu.stage_projection(n, 0.05)  # a second launch: the epoch of the progress words moves on
        cpu.projection(n, 0.05)
        cpu.projection(n, 0.05)
        assert_same(gpu, cpu, names=("u", "v"), what=f"resident T={T} rows={rows} n={n}")
        gpu.close()


@pytest.mark.parametrize("graph", [0, 1])
def test_resident_full_steps_bit_exact(graph):
    """Whole steps with the resident projection (forces folded into its load, extrapolation into its store), eager and
    as graph replays, against the oracle."""
    cfg = baseline_config(1, width=640, height=360)
    cfg["sim.projection.n"] = 11
    gpu, cpu = pair(cfg)
    gpu.set_option("resident", 2)
    gpu.set_option("use_graph", graph)
    gpu.run(3)
    assert gpu.get_option("plan_resident") == 1
    for _ in range(3):
        cpu.step(None, cfg.c.d_t)
    assert_same(gpu, cpu, what=f"resident steps graph={graph}")


def test_resident_full_size_equals_tiled_passes():
    """BASELINE configs[1] (1920x1080, n = 50): the resident plan the tuner may choose gives the bits of the multi-pass
    tiled plan over three whole steps."""
    cfg = baseline_config(1)
    u, v, sm = synthetic_fields(cfg.c.width, cfg.c.height)
    out = []
    for resident in (0, 2):
        f = Fluid(cfg)
        for name, a in (("u", u), ("v", v), ("smoke", sm)):
            f.set_field(name, a)
        f.set_option("resident", resident)
        f.run(3)
        assert f.get_option("plan_resident") == (1 if resident else 0)
        out.append({n: f.get_field(n) for n in ("u", "v", "smoke")})
        f.close()
    for n in ("u", "v", "smoke"):
        assert np.array_equal(out[0][n].view(np.uint32), out[1][n].view(np.uint32)), n


@pytest.mark.parametrize("height,cy,radius", [(470, 235, 60.0), (333, 100, 45.0), (391, 300, 80.0)])
@pytest.mark.parametrize("rows,T", [(8, 4), (12, 8), (10, 5)])
def test_listed_tiles_cut_and_shifted_change_no_bit(height, cy, radius, rows, T):
    """Whole-domain passes run over an explicit tile list: obstacle tiles cut in two along y, the last tile row moved up
    so that the bottom wall ends a warp, wall rows from the table at compile-time positions.  Same bits as the oracle
    (and hence as the regular grid), with a large disc that makes several tiles expensive."""
    cfg = baseline_config(1, width=520, height=height)
    cfg["sim.obstacle.center_x"] = 260
    cfg["sim.obstacle.center_y"] = cy
    cfg["sim.obstacle.radius"] = radius
    gpu, cpu = pair(cfg)
    gpu.set_option("resident", 0)
    gpu.set_option("temporal_block", T)
    gpu.set_option("tile_rows_per_warp", rows)
    assert gpu.get_option("split_tiles") == 1
    gpu.stage_projection(2 * T + 1, 0.05)
    cpu.projection(2 * T + 1, 0.05)
    assert_same(gpu, cpu, names=("u", "v"), what=f"listed tiles {height} {rows} {T}")
    gpu.run(2)  # and through whole steps (forces on load, extrapolation before store)
    cpu.step(None, cfg.c.d_t)
    cpu.step(None, cfg.c.d_t)
    assert_same(gpu, cpu, what="listed tiles, steps")


@pytest.mark.parametrize("width,height,radius", [(1920, 1080, 36.0), (520, 470, 60.0), (700, 333, 45.0)])
def test_tile_list_partitions_the_domain(width, height, radius):
    """Covering invariants of the explicit tile list a whole-domain pass runs over (sayal_debug_tile_list): the written
    rectangles partition the array exactly (every cell written by one tile), every tile holds its written rows plus a
    halo of 2 T rows wherever it has a neighbour, at most 16 x rows-per-warp rows, and a disc this large gets cut."""
    cfg = baseline_config(1, width=width, height=height)
    cfg["sim.obstacle.radius"] = radius
    cfg["sim.obstacle.center_x"], cfg["sim.obstacle.center_y"] = width // 2, height // 2
    gpu = Fluid(cfg)
    gpu.set_option("resident", 0)
    gpu.set_option("autotune", 0)
    T, rows_per_warp = 8, 12
    gpu.set_option("temporal_block", T)
    gpu.set_option("tile_rows_per_warp", rows_per_warp)
    gpu.stage_projection(T, 0.05)  # one pass of T iterations builds the list
    tiles = gpu.debug_tile_list(T)
    assert len(tiles) > 0
    pitch = gpu.get_option("pitch")
    halo_y, halo_x, TW, TH = 2 * T, (2 * T + 3) & ~3, 128, 16 * rows_per_warp
    cover = np.zeros((height, pitch), dtype=np.int32)
    for X0, Y0, y_end, vy0, vy1 in tiles:
        vx0 = 0 if X0 == 0 else X0 + halo_x
        vx1 = pitch if X0 + TW >= pitch else X0 + TW - halo_x
        cover[vy0:vy1, vx0:vx1] += 1
        assert Y0 <= vy0 < vy1 <= y_end <= min(Y0 + TH, height) and X0 % 4 == 0
        assert vy0 == 0 or vy0 - Y0 >= halo_y          # a tile edge with a neighbour keeps its halo
        assert vy1 == height or y_end - vy1 >= halo_y
    assert (cover == 1).all()
    regular = -(-(pitch - TW) // (TW - 2 * halo_x)) + 1
    assert len(tiles) >= regular  # at least one tile row
    base_rows = 1 if height <= TH else -(-(height - TH) // (TH - 2 * halo_y)) + 1
    assert len(tiles) > regular * base_rows  # the disc's tiles were cut in two
    gpu.close()


def test_projection_with_pressure_and_range():
    cfg = baseline_config(0)
    gpu, cpu = pair(cfg)
    gpu.stage_zero_pressure()
    cpu.zero_pressure()
    gpu.stage_projection(50, 0.05)
    cpu.projection(50, 0.05)
    assert_same(gpu, cpu, names=("u", "v", "p"), what="projection+pressure")
    assert (gpu.min_pressure, gpu.max_pressure) == cpu.pressure_range()


@pytest.mark.parametrize("rows", [8, 10, 12])
@pytest.mark.parametrize("T", [1, 3, 7, 16])
def test_tiled_projection_accumulates_pressure(T, rows):
    """enable_pressure on the tiled path: the tile's p lives in shared memory and takes one FMA per cell update
    (fluid.cu:225-226); any tile plan == the plain half-sweep kernel == the oracle, bit for bit — two projections in
    a row (p keeps accumulating across passes and calls), odd width, obstacle, d_t that is not a power of two."""
    cfg = Config.defaults(333, 210, **{"fluid.viscosity": 0.0, "sim.enable_pressure": 1, "sim.physics.g": -5.0,
                                       "sim.obstacle.radius": 21.0, "fluid.density": 1.3})
    tiled, cpu = pair(cfg)
    plain, _ = pair(cfg)
    plain.set_option("projection_kernel", 0)
    tiled.set_option("temporal_block", T)
    tiled.set_option("tile_rows_per_warp", rows)
    for f in (tiled, plain):
        f.stage_zero_pressure()
    cpu.zero_pressure()
    for n, d_t in ((2 * T + 3, 0.03), (5, 0.07)):
        tiled.stage_projection(n, d_t)
        plain.stage_projection(n, d_t)
        cpu.projection(n, d_t)
    assert_same(tiled, cpu, names=("u", "v", "p"), what=f"tiled+pressure T={T} rows={rows}")
    assert_same(plain, cpu, names=("u", "v", "p"), what="plain+pressure")
    assert (tiled.min_pressure, tiled.max_pressure) == cpu.pressure_range()


def test_extrapolation_bit_exact():
    cfg = CONFIGS["odd_203x157"]()
    gpu, cpu = pair(cfg)
    gpu.stage_extrapolation()
    cpu.extrapolation()
    assert_same(gpu, cpu, names=("u", "v"), what="extrapolation")


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_advection_bit_exact(name, kernel):
    """kernel 0 = plain per-cell kernels, 1 = shared-memory tile kernels, 2 = direct geometry-word kernels."""
    cfg = CONFIGS[name]()
    cfg["sim.smoke.enable_decay"] = 1
    cfg["sim.smoke.decay_rate"] = 0.3
    gpu, cpu = pair(cfg)
    gpu.set_option("advect_kernel", kernel)
    gpu.stage_advect_velocity(0.05)
    cpu.advect_velocity(0.05)
    assert_same(gpu, cpu, names=("u", "v"), what="velocity advection")
    gpu.stage_advect_smoke(0.05)
    cpu.advect_smoke(0.05)
    assert_same(gpu, cpu, names=("smoke",), what="smoke advection + decay")


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("amplitude", [150.0, 400.0, 3000.0])
def test_advection_fast_flow_far_backtrace(kernel, amplitude):
    """Wind-tunnel speeds and beyond: back-traces of 7 / 20 / 150 cells — inside the staged window, beyond it
    (global fallback of the tile kernels), and leaving the domain (fluid.cu:422-424)."""
    cfg = baseline_config(1, width=384, height=216)
    gpu, cpu = pair(cfg, amplitude=amplitude)
    gpu.set_option("advect_kernel", kernel)
    gpu.stage_advect_velocity(0.05)
    cpu.advect_velocity(0.05)
    gpu.stage_advect_smoke(0.05)
    cpu.advect_smoke(0.05)
    assert_same(gpu, cpu, what="fast advection")


def test_sample_velocity_bit_exact():
    cfg = CONFIGS["closed_disc_130x170"]()
    gpu, cpu = pair(cfg)
    rng = np.random.default_rng(5)
    xs = rng.uniform(-5, 135, 4096).astype(np.float32)
    ys = rng.uniform(-5, 175, 4096).astype(np.float32)
    xs[:8] = [0, 1, 0.5, 64.5, 129.5, 130, 65, 64.999]
    ys[:8] = [0, 1, 0.5, 100.5, 0.5, 170, 120, 85.5]
    gu, gv = gpu.get_general_velocity(xs, ys)
    cu, cv = cpu.sample_velocity(xs, ys)
    assert np.array_equal(gu, cu) and np.array_equal(gv, cv)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_steps_bit_exact(name):
    cfg = CONFIGS[name]()
    cfg["sim.projection.n"] = 20
    gpu, cpu = pair(cfg)
    names = ("u", "v", "smoke") + (("p",) if cfg.c.enable_pressure else ())
    for step in range(4):
        gpu.update(None, cfg.c.d_t)
        cpu.step(None, cfg.c.d_t)
        assert_same(gpu, cpu, names=names, what=f"step {step}")
    if cfg.c.enable_pressure:
        assert (gpu.min_pressure, gpu.max_pressure) == cpu.pressure_range()


@pytest.mark.parametrize("width,height", [(203, 157), (384, 216), (517, 301)])
def test_forces_folded_into_first_projection_pass(width, height):
    """fuse_forces (default): v += g d_t and the inlet are applied while the step's first tiled pass loads its
    tile, and the boundary extrapolation while the last pass stores.  Same bits as the separate kernels and as the
    oracle — with gravity, an inlet wider than one lane's four columns, several smoke bands, and widths with
    W % 4 = 3, 0, 1 (column W-2 in the same lane as W-1, or in the previous one)."""
    cfg = Config.defaults(width, height, **{"fluid.viscosity": 0.0, "sim.physics.g": -4.0, "sim.projection.n": 13,
                                            "sim.wind_tunnel.speed": 80.0, "sim.wind_tunnel.smoke_length": 6,
                                            "sim.wind_tunnel.smoke_count": 3, "sim.wind_tunnel.smoke_height": 5,
                                            "sim.wind_tunnel.pipe_height": height // 3})
    fused, cpu = pair(cfg)
    plain, _ = pair(cfg)
    plain.set_option("fuse_forces", 0)
    plain.set_option("fuse_extrapolation", 0)
    fused.set_option("fuse_forces", 1)
    fused.set_option("fuse_extrapolation", 1)
    for step in range(3):
        l0, l1 = fused.launch_count, plain.launch_count
        fused.update(None, cfg.c.d_t)
        plain.update(None, cfg.c.d_t)
        cpu.step(None, cfg.c.d_t)
        assert_same(fused, cpu, what=f"fused, step {step}")
        assert_same(plain, cpu, what=f"separate kernel, step {step}")
        assert (plain.launch_count - l1) - (fused.launch_count - l0) == 2  # forces and extrapolation kernels are gone
    fused.run(2)
    plain.run(2)
    for _ in range(2):
        cpu.step(None, cfg.c.d_t)
    assert_same(fused, cpu, what="fused, graph replay")
    assert_same(plain, cpu, what="separate kernel, graph replay")


def test_interactive_source_steps():
    cfg = baseline_config(0)
    cfg["sim.projection.n"] = 10
    gpu, cpu = pair(cfg)
    for step in range(3):
        src = Source(True, 1.0, 3.0, (100 + 10 * step, 70))
        gpu.update(src, 0.04)
        cpu.step(src._c(), 0.04)
    assert_same(gpu, cpu, names=("u", "v", "smoke", "p"), what="interactive")


@pytest.mark.parametrize("graph", [0, 1])
def test_run_equals_repeated_update(graph):
    cfg = baseline_config(1, width=384, height=216)
    cfg["sim.projection.n"] = 11  # odd number of passes: exercises the buffer-parity bookkeeping
    a, _ = pair(cfg)
    b, _ = pair(cfg)
    a.set_option("use_graph", graph)
    a.run(7)
    a.sync()
    for _ in range(7):
        b.update(None)
    for n in ("u", "v", "smoke"):
        assert np.array_equal(a.get_field(n), b.get_field(n)), n
    a.run(4)  # cached graph, different starting parity
    for _ in range(4):
        b.update(None)
    assert np.array_equal(a.get_field("u"), b.get_field("u"))
    assert a.launch_count == b.launch_count


@pytest.mark.parametrize("width,T", [(352, 3), (352, 4), (384, 15), (384, 16), (336, 5), (336, 6), (1920, 4)])
def test_tile_ending_exactly_on_the_last_column(width, T):
    """Regression: when 128 + k*stride == W the last tile has no right halo; its carrier column is the real
    border column and must not take the all-open path (found by the slab test at T=16)."""
    cfg = baseline_config(1, width=width, height=200)
    gpu, cpu = pair(cfg)
    gpu.set_option("temporal_block", T)
    gpu.stage_projection(2 * T + 1, 0.05)
    cpu.projection(2 * T + 1, 0.05)
    assert_same(gpu, cpu, names=("u", "v"), what=f"exact fit W={width} T={T}")


@pytest.mark.parametrize("graph", [0, 1])
def test_tile_issue_order_changes_no_bit(graph):
    """order_tiles (default): tiles with walls / obstacle rims are issued first.  Any order gives the same bits."""
    cfg = baseline_config(1, width=1000, height=700)
    cfg["sim.projection.n"] = 9
    a, cpu = pair(cfg)
    b, _ = pair(cfg)
    b.set_option("order_tiles", 0)
    for f in (a, b):
        f.set_option("use_graph", graph)
        f.run(3)
    for _ in range(3):
        cpu.step(None, cfg.c.d_t)
    assert_same(a, cpu, what="ordered tiles")
    assert_same(b, cpu, what="row-major tiles")


def test_autotuned_plan_is_invisible():
    """The autotuner times candidate plans on the live arrays; state, results and launch count must not show it."""
    cfg = baseline_config(1, width=640, height=400)
    a, cpu = pair(cfg)
    b, _ = pair(cfg)
    b.set_option("autotune", 0)
    a.stage_projection(50, 0.05)
    b.stage_projection(50, 0.05)
    cpu.projection(50, 0.05)
    assert_same(a, cpu, names=("u", "v"), what="autotuned")
    assert_same(b, cpu, names=("u", "v"), what="model plan")
    assert a.get_option("plan_temporal_block") >= 1 and a.get_option("plan_rows_per_warp") in (8, 10, 12)


def test_cell_size_2_generic_path():
    """Integer cell_size != 1 takes the generic (HC = 0) advection kernels and the hf pressure multiply."""
    cfg = baseline_config(0, width=160, height=120)
    cfg["sim.cell_size"] = 2.0
    cfg["sim.projection.n"] = 6
    gpu, cpu = pair(cfg)
    for _ in range(2):
        gpu.update(None)
        cpu.step(None)
    assert_same(gpu, cpu, names=("u", "v", "smoke", "p"), what="cell_size 2")


def test_zero_start_wind_tunnel_matches_oracle():
    """The reference's natural all-zero start (fluid.cu:104-109): inlet drives everything."""
    cfg = baseline_config(1, width=384, height=216)
    gpu, cpu = Fluid(cfg), OracleSim(cfg.c)
    for _ in range(5):
        gpu.update(None)
        cpu.step(None)
    assert_same(gpu, cpu, what="zero start")
    assert gpu.get_field("smoke").max() > 0


def test_full_size_1920x1080_one_step_and_properties():
    """BASELINE config 2 at full size: one step bit-exact vs the oracle (seconds on one core); then
    size-independent properties: tiled == plain, and projection contracts the divergence."""
    cfg = baseline_config(1)
    gpu, cpu = pair(cfg)
    gpu.update(None)
    cpu.step(None)
    assert_same(gpu, cpu, what="1920x1080 step")


def test_tile_advection_equals_plain_at_full_size():
    """Size-independent property at BASELINE sizes: the tile kernels and the plain kernels agree bit for bit
    (1920x1080 wind tunnel after two steps, so the inlet jet and the wake are in the field)."""
    cfg = baseline_config(1)
    u, v, sm = synthetic_fields(1920, 1080)
    outs = []
    for kernel in (0, 1, 2):
        f = Fluid(cfg)
        f.set_option("advect_kernel", kernel)
        for name, a in (("u", u), ("v", v), ("smoke", sm)):
            f.set_field(name, a)
        f.run(2)
        outs.append([f.get_field(n) for n in ("u", "v", "smoke")])
        f.close()
    for other in outs[1:]:
        for a, b, n in zip(outs[0], other, ("u", "v", "smoke")):
            assert np.array_equal(a, b), n


def test_large_grid_properties_3840x2160():
    cfg = baseline_config(2)
    u, v, sm = synthetic_fields(3840, 2160)
    outs = []
    for kernel in (0, 1):
        f = Fluid(cfg)
        f.set_field("u", u)
        f.set_field("v", v)
        f.set_option("projection_kernel", kernel)
        f.stage_projection(30, 0.05)
        outs.append((f.get_field("u"), f.get_field("v"), f.is_solid))
        f.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    uu, vv, solid = outs[1]

    def mean_div(a, b):
        d = (a[1:-1, 2:] - a[1:-1, 1:-1]) + (b[:-2, 1:-1] - b[1:-1, 1:-1])
        return float(np.abs(d[:, :-1][solid[1:-1, 1:-2] == 0]).mean())

    assert mean_div(uu, vv) < 0.25 * mean_div(u, v)


def test_batched_and_device_field_io_and_the_stream_gate():
    """sayal_set_fields / sayal_get_fields (batched, host), sayal_set_field_device / sayal_get_field_device (device
    buffers in the reference layout, odd width: padded pitch inside) move the same bits as the one-field calls; steps
    enqueued behind sayal_stream_hold run when sayal_stream_release says so and give the same result."""
    import torch
    cfg = Config.defaults(203, 157, **{"fluid.viscosity": 0.0, "sim.wind_tunnel.speed": 60.0})
    W, H = cfg.c.width, cfg.c.height
    u, v, sm = synthetic_fields(W, H)
    a, b = Fluid(cfg), Fluid(cfg)
    for n, x in (("u", u), ("v", v), ("smoke", sm)):
        a.set_field(n, x)
    pinned = {n: torch.from_numpy(x.copy()).pin_memory() for n, x in (("u", u), ("v", v), ("smoke", sm))}
    b.set_fields_from({n: t.data_ptr() for n, t in pinned.items()})
    b.sync()
    for n in ("u", "v", "smoke"):
        assert np.array_equal(a.get_field(n), b.get_field(n)), n
    a.run(2)
    b.stream_hold()          # the two steps wait on the device ...
    b.run(2)
    b.stream_release()       # ... until the host opens the gate
    out = {n: torch.empty((H, W), dtype=torch.float32).pin_memory() for n in ("u", "v", "smoke")}
    b.get_fields_into({n: t.data_ptr() for n, t in out.items()})
    for n in ("u", "v", "smoke"):
        assert np.array_equal(a.get_field(n), out[n].numpy()), n
    dev = torch.empty((H, W), dtype=torch.float32, device="cuda")
    a.get_field_device("smoke", dev.data_ptr())
    a.sync()
    assert np.array_equal(dev.cpu().numpy(), a.get_field("smoke"))
    dev.mul_(0.5)
    torch.cuda.synchronize()
    a.set_field_device("smoke", dev.data_ptr())
    a.sync()
    assert np.array_equal(a.get_field("smoke"), dev.cpu().numpy())
    a.close()
    b.close()


def test_errors_are_codes_not_exits():
    from opensayal_b200 import SayalError
    cfg = Config.defaults(2, 2)
    with pytest.raises(SayalError):
        Fluid(cfg)
    cfg = Config.defaults(64, 64)
    with pytest.raises(SayalError):
        Fluid(cfg, device=99)
    f = Fluid(cfg)
    with pytest.raises(SayalError):
        f.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        f.set_field("u", np.zeros((3, 3), np.float32))
