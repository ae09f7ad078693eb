"""Synthetic initial fields and the BASELINE.json configurations.

Host-side data generation only (numpy); nothing here is on the step path.
"""
from __future__ import annotations

import numpy as np

from .fluid import Config

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64(idx: np.ndarray, seed: int) -> np.ndarray:
    """splitmix64 of (seed, idx) -> uint64; wraps mod 2^64 like the C reference implementation."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) * _GOLDEN + np.uint64(seed) * _GOLDEN + _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def noise(shape, seed: int) -> np.ndarray:
    """Deterministic uniform noise in [-1, 1), float32."""
    n = int(np.prod(shape))
    bits = _splitmix64(np.arange(n, dtype=np.uint64), seed) >> np.uint64(40)  # 24 random bits
    return (bits.astype(np.float32) * np.float32(2.0 ** -23) - np.float32(1.0)).reshape(shape)


def synthetic_fields(width: int, height: int, seed: int = 1234, amplitude: float = 40.0, rows=None):
    """u, v, smoke in the reference layout (row r is j = H-1-r).  `rows=(r0, n)` generates only memory rows
    [r0, r0+n) of the same global field (for y-slabs)."""
    r0, n = (0, height) if rows is None else rows
    r = np.arange(r0, r0 + n, dtype=np.float64)[:, None]
    y = (height - 1 - r) + 0.5
    x = np.arange(width, dtype=np.float64)[None, :] + 0.5
    two_pi = 2.0 * np.pi
    idx0 = r0 * width
    shape = (n, width)

    def nz(k):
        cells = int(np.prod(shape))
        bits = _splitmix64(np.arange(idx0, idx0 + cells, dtype=np.uint64), seed + k) >> np.uint64(40)
        return (bits.astype(np.float32) * np.float32(2.0 ** -23) - np.float32(1.0)).reshape(shape)

    u = amplitude * np.sin(two_pi * 3 * x / width) * np.cos(two_pi * 2 * y / height) + 5.0 * nz(1)
    v = -amplitude * np.cos(two_pi * 3 * x / width) * np.sin(two_pi * 2 * y / height) + 5.0 * nz(2)
    smoke = 0.5 * (1.0 + np.sin(two_pi * 8 * x / width) * np.sin(two_pi * 8 * y / height))
    return u.astype(np.float32), v.astype(np.float32), smoke.astype(np.float32)


def baseline_config(index: int, width: int | None = None, height: int | None = None, defaults=None) -> Config:
    """BASELINE.json `configs[index]` restated as concrete inputs.  All of them:
    cell_size 1, density 1, drag 0, viscosity 0 (H1), o 1.9, d_t 0.05, inactive source.
    defaults: callable (width, height) -> SayalConfig replacing the library's sayal_config_defaults (the reference
    arm of bench.py passes oracle.reference_defaults so that it never maps the product library)."""
    common = {"fluid.viscosity": 0.0, "fluid.drag_coeff": 0.0, "sim.projection.o": 1.9, "sim.time.d_t": 0.05}
    if defaults is not None:
        common["_struct"] = defaults
    if index == 0:  # 256x144 gravity tank
        w, h = width or 256, height or 144
        return Config.defaults(w, h, **common, **{
            "sim.physics.g": -5.0, "sim.enable_drain": 0, "sim.enable_pressure": 1, "sim.enable_smoke": 1,
            "sim.wind_tunnel.speed": 0.0, "sim.wind_tunnel.smoke": 1.0, "sim.wind_tunnel.smoke_length": 0,
            "sim.wind_tunnel.pipe_height": h // 4, "sim.obstacle.enable": 0, "sim.projection.n": 50})
    if index in (1, 2, 3, 4):
        dims = {1: (1920, 1080), 2: (3840, 2160), 3: (16384, 16384), 4: (16384, 8192)}[index]
        w, h = width or dims[0], height or dims[1]
        n = {1: 50, 2: 100, 3: 50, 4: 200}[index]
        scale = h / 1080.0 if index in (1, 2) else None
        ph = int(270 * scale) if scale else h // 4
        radius = 36.0 * scale if scale else float(min(w, h) // 30)
        over = {
            "sim.physics.g": 0.0, "sim.enable_drain": 1, "sim.enable_pressure": 0, "sim.enable_smoke": 1,
            "sim.wind_tunnel.speed": 200.0, "sim.wind_tunnel.smoke": 1.0, "sim.wind_tunnel.smoke_length": 1,
            "sim.wind_tunnel.pipe_height": ph, "sim.obstacle.enable": 1, "sim.obstacle.center_x": w // 2,
            "sim.obstacle.center_y": h // 2, "sim.obstacle.radius": radius, "sim.projection.n": n}
        if index == 2:
            over.update({"sim.smoke.enable_decay": 1, "sim.smoke.decay_rate": 0.05})
        return Config.defaults(w, h, **common, **over)
    raise ValueError("config index 0..4")
