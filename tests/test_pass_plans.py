"""Host-side arithmetic of the tiled projection (no GPU): how a projection call is split into passes, which rows a
pass sweeps, and that the parts the tiles of a pass write partition those rows x columns exactly.  The kernel writes,
for tile (a, b), the columns / rows at least one halo away from every tile edge that has a neighbour
(projection_pack.cu, `vx0 .. vy1`); a gap would leave cells un-projected, an overlap would be a write race."""
import ctypes as C

import numpy as np
import pytest

from opensayal_b200 import load


def pass_plans(pitch, rows, own_lo, own_hi, rows_per_warp, T, n, depth):
    lib = load()
    out = (C.c_int32 * (11 * 64))()
    cnt = C.c_int32()
    rc = lib.sayal_debug_pass_plans(pitch, rows, own_lo, own_hi, rows_per_warp, T, n, depth, out, 64, C.byref(cnt))
    assert rc == 0, lib.sayal_last_error()
    keys = ("iterations", "row_lo", "row_hi", "halo_x", "halo_y", "stride_x", "stride_y", "tiles_x", "tiles_y", "tile_w",
            "tile_h")
    return [dict(zip(keys, out[11 * k: 11 * k + 11])) for k in range(cnt.value)]


def written_ranges(origin, extent_end, tile, stride, halo, tiles):
    """[lo, hi) each tile writes along one axis, as the kernel computes it."""
    out = []
    for t in range(tiles):
        p0 = origin + t * stride
        lo = origin if p0 == origin else p0 + halo
        hi = extent_end if p0 + tile >= extent_end else p0 + tile - halo
        out.append((lo, hi))
    return out


@pytest.mark.parametrize("rows_per_warp", [8, 10, 12])
@pytest.mark.parametrize("pitch,rows", [(1920, 1080), (204, 157), (128, 40), (132, 4096), (16384, 2284), (4, 4)])
def test_tiles_of_a_pass_partition_the_window(pitch, rows, rows_per_warp):
    for T in (1, 2, 3, 5, 8, 11, 16):
        plans = pass_plans(pitch, rows, 0, rows, rows_per_warp, T, 2 * T + 1, -1)
        if not plans:  # temporal block too deep for this tile height: the planner never proposes it
            continue
        for p in plans:
            assert p["halo_y"] == 2 * p["iterations"] and p["halo_x"] >= 2 * p["iterations"] and p["halo_x"] % 4 == 0
            for origin, end, tile, stride, halo, tiles in (
                    (0, pitch, p["tile_w"], p["stride_x"], p["halo_x"], p["tiles_x"]),
                    (p["row_lo"], p["row_hi"], p["tile_h"], p["stride_y"], p["halo_y"], p["tiles_y"])):
                cover = np.zeros(end, dtype=np.int32)
                for lo, hi in written_ranges(origin, end, tile, stride, halo, tiles):
                    assert origin <= lo < hi <= end, "every tile writes something, inside the window"
                    cover[lo:hi] += 1
                assert (cover[origin:end] == 1).all(), "written parts partition the window: no gap, no overlap"


@pytest.mark.parametrize("n,T", [(50, 10), (50, 8), (25, 10), (7, 16), (200, 9), (1, 1), (16, 16), (17, 16)])
def test_passes_split_evenly(n, T):
    plans = pass_plans(1920, 1080, 0, 1080, 8, T, n, -1)
    its = [p["iterations"] for p in plans]
    assert sum(its) == n and len(its) == -(-n // T)
    assert max(its) <= T and max(its) - min(its) <= 1 and its == sorted(its, reverse=True)


@pytest.mark.parametrize("halo,n", [(118, 50), (50, 25), (34, 17), (18, 9), (60, 3)])
def test_slab_window_shrinks_with_the_valid_ghost_rows(halo, n):
    """A linked slab's pass sweeps the owned rows plus the ghost rows that are still exact before it: two rows fewer
    per iteration already done, clipped to the rows held; after the call, halo - 2 n rows are still exact."""
    own_lo, own_hi, rows = halo, halo + 1080, 1080 + 2 * halo
    for T in (4, 9, 16):
        plans = pass_plans(1920, rows, own_lo, own_hi, 8, T, n, halo)
        done = 0
        for p in plans:
            depth = halo - 2 * done
            assert depth >= 2 * p["iterations"], "a pass never needs more exact ghost rows than are left"
            assert (p["row_lo"], p["row_hi"]) == (own_lo - depth, own_hi + depth)
            done += p["iterations"]
        assert done == n and halo - 2 * done >= 0
    # first / last slab: no ghost rows on the outer side
    first = pass_plans(1920, 1080 + halo, 0, 1080, 8, 8, n, halo)
    assert all(p["row_lo"] == 0 for p in first) and first[0]["row_hi"] == 1080 + halo
    last = pass_plans(1920, 1080 + halo, halo, 1080 + halo, 8, 8, n, halo)
    assert all(p["row_hi"] == 1080 + halo for p in last) and last[0]["row_lo"] == 0


def test_bad_arguments_are_errors():
    lib = load()
    out = (C.c_int32 * 11)()
    cnt = C.c_int32()
    assert lib.sayal_debug_pass_plans(1922, 1080, 0, 1080, 8, 8, 50, -1, out, 1, C.byref(cnt)) < 0   # pitch % 4
    assert lib.sayal_debug_pass_plans(1920, 1080, 0, 1080, 9, 8, 50, -1, out, 1, C.byref(cnt)) < 0   # rows per warp
    assert lib.sayal_debug_pass_plans(1920, 1080, 0, 1080, 8, 17, 50, -1, out, 1, C.byref(cnt)) < 0  # T > 16
