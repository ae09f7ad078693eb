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
    out = (C.c_int32 * (15 * 64))()
    cnt = C.c_int32()
    rc = lib.sayal_debug_pass_plans(pitch, rows, own_lo, own_hi, rows_per_warp, T, n, depth, out, 64, C.byref(cnt))
    assert rc == 0, lib.sayal_last_error()
    keys = ("iterations", "row_lo", "row_hi", "halo_x", "halo_y", "stride_x", "stride_y", "tiles_x", "tiles_y", "tile_w",
            "tile_h", "write_lo", "write_hi", "pushers0", "pushers1")
    return [dict(zip(keys, out[15 * k: 15 * k + 15])) for k in range(cnt.value)]


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


@pytest.mark.parametrize("rows_per_warp", [8, 10, 12])
@pytest.mark.parametrize("own,halo,n,T", [(1080, 18, 50, 8), (1080, 16, 50, 7), (2048, 18, 50, 8), (100, 20, 9, 5),
                                          (18, 18, 50, 8), (300, 24, 200, 8), (64, 32, 16, 16)])
def test_push_mode_passes_write_the_owned_rows_once_and_count_their_pushers(own, halo, n, T, rows_per_warp):
    """Push mode (linked slabs): pass k sweeps the owned rows +- 2 it on the sides that have a neighbour, the tiles
    write exactly the owned rows (the ghost rows of the output belong to the neighbours' pushes), and the number of
    tiles the kernel counts down before it publishes a side's flag equals the number of tiles whose written rows meet
    that side's edge band [own_lo, own_lo + halo) / [own_hi - halo, own_hi) — recomputed here from the geometry."""
    pitch = 1920
    for where in ("interior", "first", "last"):
        own_lo = 0 if where == "first" else halo
        own_hi = own_lo + own
        rows = own_hi + (0 if where == "last" else halo)
        plans = pass_plans(pitch, rows, own_lo, own_hi, rows_per_warp, T, n, -halo)
        if not plans:
            continue
        assert sum(p["iterations"] for p in plans) == n
        for p in plans:
            it = p["iterations"]
            assert 2 * it <= halo
            assert p["row_lo"] == (own_lo - 2 * it if own_lo > 0 else 0)
            assert p["row_hi"] == (own_hi + 2 * it if own_hi < rows else rows)
            assert (p["write_lo"], p["write_hi"]) == (own_lo, own_hi)
            cover = np.zeros(rows, dtype=np.int32)
            pushers = [0, 0]
            band = [(own_lo, own_lo + halo), (own_hi - halo, own_hi)]
            for lo, hi in written_ranges(p["row_lo"], p["row_hi"], p["tile_h"], p["stride_y"], p["halo_y"], p["tiles_y"]):
                lo, hi = max(lo, p["write_lo"]), min(hi, p["write_hi"])
                if lo < hi:
                    cover[lo:hi] += 1
                for d in (0, 1):
                    if lo < band[d][1] and hi > band[d][0]:
                        pushers[d] += p["tiles_x"]
            assert (cover[own_lo:own_hi] == 1).all() and cover[:own_lo].sum() == 0 and cover[own_hi:].sum() == 0
            assert p["pushers0"] == (pushers[0] if own_lo > 0 else 0)
            assert p["pushers1"] == (pushers[1] if own_hi < rows else 0)
            if own_lo > 0:
                assert p["pushers0"] >= p["tiles_x"]
            # every row of an edge band is written by some tile, so a pushing tile row always exists
            # exactness: a written row is at least 2 it rows inside the window on every side that has a neighbour
            if own_lo > 0:
                assert p["write_lo"] - p["row_lo"] >= 2 * it
            if own_hi < rows:
                assert p["row_hi"] - p["write_hi"] >= 2 * it


def test_push_mode_rejects_a_pass_deeper_than_the_halo():
    lib = load()
    out = (C.c_int32 * (15 * 64))()
    cnt = C.c_int32()
    # T = 10 needs 20 ghost rows per pass, the slab has 18
    assert lib.sayal_debug_pass_plans(1920, 1080 + 36, 18, 1098, 8, 10, 50, -18, out, 64, C.byref(cnt)) < 0


def test_bad_arguments_are_errors():
    lib = load()
    out = (C.c_int32 * 15)()
    cnt = C.c_int32()
    assert lib.sayal_debug_pass_plans(1922, 1080, 0, 1080, 8, 8, 50, -1, out, 1, C.byref(cnt)) < 0   # pitch % 4
    assert lib.sayal_debug_pass_plans(1920, 1080, 0, 1080, 9, 8, 50, -1, out, 1, C.byref(cnt)) < 0   # rows per warp
    assert lib.sayal_debug_pass_plans(1920, 1080, 0, 1080, 8, 17, 50, -1, out, 1, C.byref(cnt)) < 0  # T > 16
