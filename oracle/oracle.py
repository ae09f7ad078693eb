"""ctypes wrappers of the two checkers: liboracle.so (CPU restatement) and _ref/libsayal_ref.so (the
reference's own CUDA code).  TEST INFRASTRUCTURE — imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference legs, never by opensayal_b200.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from opensayal_b200._abi import ARROW_DTYPE, FIELD_NAMES, IS_SOLID, TOTAL_S, SayalConfig, SayalSource, SayalVisual

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "liboracle.so"
REF_LIB = HERE / "_ref" / "libsayal_ref.so"


def reference_defaults(width: int = 1920, height: int = 1080) -> SayalConfig:
    """The defaults of the reference's ConfigParser::parse (/root/reference/src/config_parser.cpp:21-119) restated
    in Python, so that the reference arm of bench.py can build its configuration without mapping the product library
    (sayal_config_defaults).  tests/test_config.py checks that the two agree."""
    c = SayalConfig()
    c.width, c.height, c.cell_size = width, height, 1.0
    c.enable_drain, c.enable_pressure, c.enable_smoke, c.enable_interactive = 1, 0, 1, 0
    c.proj_n, c.proj_o = 50, 1.9
    c.wt_pipe_height, c.wt_pipe_length, c.wt_smoke_length = height // 4, 0, 1
    c.wt_smoke_height, c.wt_smoke_count, c.wt_speed, c.wt_smoke = height // 4, 1, 0.0, 1.0
    c.g, c.d_t, c.enable_real_time, c.real_time_multiplier = 0.0, 0.05, 0, 1.0
    c.smoke_enable_decay, c.smoke_decay_rate = 0, 0.05
    c.obstacle_enable, c.obstacle_center_x, c.obstacle_center_y = 1, width // 2, height // 2
    c.obstacle_radius = min(height, width) / 30.0
    c.density, c.drag_coeff, c.viscosity = 1.0, 0.0, 0.001
    c.block_size_x, c.block_size_y = 64, 1
    return c


def build(ref: bool = True) -> None:
    subprocess.run(["make", "-C", str(HERE), "liboracle.so"] + (["ref"] if ref else []), check=True,
                   capture_output=True, text=True)


def _load_oracle():
    if not ORACLE_LIB.exists():
        build(ref=False)
    lib = C.CDLL(str(ORACLE_LIB))
    vp, cfgp, srcp = C.c_void_p, C.POINTER(SayalConfig), C.POINTER(SayalSource)
    sig = {
        "oracle_build_masks": (None, [cfgp, vp, vp]),
        "oracle_create": (vp, [cfgp]),
        "oracle_destroy": (None, [vp]),
        "oracle_set_threads": (None, [vp, C.c_int]),
        "oracle_field": (vp, [vp, C.c_int]),
        "oracle_forces": (None, [vp, srcp, C.c_float]),
        "oracle_zero_pressure": (None, [vp]),
        "oracle_projection": (None, [vp, C.c_int, C.c_float]),
        "oracle_pressure_range": (None, [vp]),
        "oracle_get_pressure_range": (None, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "oracle_extrapolation": (None, [vp]),
        "oracle_advect_velocity": (None, [vp, C.c_float]),
        "oracle_advect_smoke": (None, [vp, C.c_float]),
        "oracle_decay_smoke": (None, [vp, C.c_float]),
        "oracle_step": (None, [vp, srcp, C.c_float]),
        "oracle_sample_velocity": (None, [vp, C.c_int, vp, vp, vp, vp]),
        "oracle_render_pixels": (None, [vp, vp]),
        "oracle_diffusion": (None, [vp, C.c_int, C.c_float]),
        "oracle_path_lines": (None, [vp, C.POINTER(SayalVisual), C.c_float, vp, vp]),
        "oracle_arrows": (None, [vp, C.POINTER(SayalVisual), vp]),
        "oracle_max_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = _load_oracle()
    return _oracle


def build_masks(cfg: SayalConfig):
    """(is_solid, total_s) int32 arrays, shape (H, W), reference layout."""
    solid = np.empty((cfg.height, cfg.width), np.int32)
    total = np.empty((cfg.height, cfg.width), np.int32)
    oracle_lib().oracle_build_masks(C.byref(cfg), solid.ctypes.data, total.ctypes.data)
    return solid, total


class OracleSim:
    """CPU restatement of class Fluid; fields are numpy views onto the C arrays (they move on swap)."""

    def __init__(self, cfg: SayalConfig, threads: int = 1):
        self.lib = oracle_lib()
        self.cfg = cfg.copy()
        self.h = self.lib.oracle_create(C.byref(self.cfg))
        if not self.h:
            raise MemoryError("oracle_create failed")
        self.lib.oracle_set_threads(self.h, threads)
        self.shape = (cfg.height, cfg.width)

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def field(self, name: str) -> np.ndarray:
        fid = FIELD_NAMES[name]
        ptr = self.lib.oracle_field(self.h, fid)
        ctype = C.c_int32 if fid in (IS_SOLID, TOTAL_S) else C.c_float
        n = self.shape[0] * self.shape[1]
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(self.shape)

    def get_field(self, name: str) -> np.ndarray:
        return self.field(name).copy()

    def set_field(self, name: str, a: np.ndarray) -> None:
        self.field(name)[...] = a

    @staticmethod
    def _src(source):
        return C.byref(source) if source is not None else None

    def forces(self, source, d_t):
        self.lib.oracle_forces(self.h, self._src(source), d_t)

    def zero_pressure(self):
        self.lib.oracle_zero_pressure(self.h)

    def projection(self, iterations, d_t):
        self.lib.oracle_projection(self.h, iterations, d_t)

    def pressure_range(self):
        self.lib.oracle_pressure_range(self.h)
        mn, mx = C.c_float(), C.c_float()
        self.lib.oracle_get_pressure_range(self.h, C.byref(mn), C.byref(mx))
        return mn.value, mx.value

    def extrapolation(self):
        self.lib.oracle_extrapolation(self.h)

    def advect_velocity(self, d_t):
        self.lib.oracle_advect_velocity(self.h, d_t)

    def advect_smoke(self, d_t):
        self.lib.oracle_advect_smoke(self.h, d_t)
        self.lib.oracle_decay_smoke(self.h, d_t)

    def step(self, source=None, d_t=None):
        self.lib.oracle_step(self.h, self._src(source), self.cfg.d_t if d_t is None else d_t)

    def render_pixels(self, prefill: int = 0) -> np.ndarray:
        """RGBA8888 frame (H, W) uint32 of the current state (graphics_handler.cu:258-302)."""
        out = np.full(self.shape, prefill, np.uint32)
        self.lib.oracle_render_pixels(self.h, out.ctypes.data)
        return out

    def diffusion(self, iterations, d_t):
        self.lib.oracle_diffusion(self.h, iterations, d_t)

    def path_lines(self, visual: SayalVisual, d_t=None):
        """update_traces over Fluid::trace: (xs, ys), shape (n_y, n_x, length)."""
        ny, nx = self.shape[0] // visual.path_line_distance, self.shape[1] // visual.path_line_distance
        xs = np.zeros((ny, nx, visual.path_line_length), np.int32)
        ys = np.zeros_like(xs)
        self.lib.oracle_path_lines(self.h, C.byref(visual), self.cfg.d_t if d_t is None else d_t, xs.ctypes.data,
                                   ys.ctypes.data)
        return xs, ys

    def arrows(self, visual: SayalVisual) -> np.ndarray:
        ny, nx = self.shape[0] // visual.arrows_distance, self.shape[1] // visual.arrows_distance
        out = np.zeros((ny, nx), dtype=ARROW_DTYPE)
        self.lib.oracle_arrows(self.h, C.byref(visual), out.ctypes.data)
        return out

    def sample_velocity(self, xs, ys):
        xs = np.ascontiguousarray(xs, np.float32)
        ys = np.ascontiguousarray(ys, np.float32)
        ou, ov = np.empty_like(xs), np.empty_like(ys)
        self.lib.oracle_sample_velocity(self.h, xs.size, xs.ctypes.data, ys.ctypes.data, ou.ctypes.data, ov.ctypes.data)
        return ou, ov


class RefSim:
    """The reference's own `Fluid` (CUDA, fast-math, sm_100) behind oracle/ref_shim.cu.  GPU only."""

    def __init__(self, cfg: SayalConfig, device: int = 0):
        if not REF_LIB.exists():
            raise FileNotFoundError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        lib = C.CDLL(str(REF_LIB))
        vp, cfgp, srcp = C.c_void_p, C.POINTER(SayalConfig), C.POINTER(SayalSource)
        lib.ref_create.restype, lib.ref_create.argtypes = C.c_int, [cfgp, C.c_int, C.POINTER(vp)]
        lib.ref_destroy.restype, lib.ref_destroy.argtypes = None, [vp]
        lib.ref_step.restype, lib.ref_step.argtypes = C.c_int, [vp, srcp, C.c_float]
        lib.ref_run_timed.restype, lib.ref_run_timed.argtypes = C.c_int, [vp, C.c_int, C.c_float, C.POINTER(C.c_float)]
        lib.ref_get_field.restype, lib.ref_get_field.argtypes = C.c_int, [vp, C.c_int, vp]
        lib.ref_set_field.restype, lib.ref_set_field.argtypes = C.c_int, [vp, C.c_int, vp]
        lib.ref_pressure_range.restype, lib.ref_pressure_range.argtypes = C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self.lib = lib
        self.cfg = cfg.copy()
        self.h = vp()
        rc = lib.ref_create(C.byref(self.cfg), device, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"ref_create failed ({rc})")
        self.shape = (cfg.height, cfg.width)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_field(self, name: str) -> np.ndarray:
        fid = FIELD_NAMES[name]
        out = np.empty(self.shape, np.int32 if fid in (IS_SOLID, TOTAL_S) else np.float32)
        assert self.lib.ref_get_field(self.h, fid, out.ctypes.data) == 0
        return out

    def set_field(self, name: str, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, np.float32)
        assert self.lib.ref_set_field(self.h, FIELD_NAMES[name], a.ctypes.data) == 0

    def step(self, source=None, d_t=None):
        rc = self.lib.ref_step(self.h, C.byref(source) if source is not None else None,
                               self.cfg.d_t if d_t is None else d_t)
        assert rc == 0

    def run_timed(self, steps: int, d_t=None) -> float:
        ms = C.c_float()
        assert self.lib.ref_run_timed(self.h, steps, self.cfg.d_t if d_t is None else d_t, C.byref(ms)) == 0
        return ms.value

    def pressure_range(self):
        mn, mx = C.c_float(), C.c_float()
        self.lib.ref_pressure_range(self.h, C.byref(mn), C.byref(mx))
        return mn.value, mx.value


class RefGfx:
    """The reference's own GraphicsHandler (graphics_handler.cu, unmodified) run headless over recording SDL stubs
    (oracle/ref_shim.cu): update() is graphics.update(fluid, d_t) of main.cu:98; what it handed to SDL is returned."""

    def __init__(self, ref: RefSim, visual: SayalVisual):
        lib = ref.lib
        vp = C.c_void_p
        lib.ref_gfx_create.restype, lib.ref_gfx_create.argtypes = C.c_int, [C.POINTER(SayalConfig), C.POINTER(SayalVisual), C.POINTER(vp)]
        lib.ref_gfx_destroy.restype, lib.ref_gfx_destroy.argtypes = None, [vp]
        lib.ref_gfx_update.restype, lib.ref_gfx_update.argtypes = C.c_int, [vp, vp, C.c_float]
        lib.ref_gfx_pixels.restype, lib.ref_gfx_pixels.argtypes = C.c_int64, [vp, C.c_int64]
        lib.ref_gfx_polylines.restype, lib.ref_gfx_polylines.argtypes = C.c_int64, [vp, C.c_int64, C.POINTER(C.c_int32)]
        lib.ref_gfx_segments.restype, lib.ref_gfx_segments.argtypes = C.c_int64, [vp, C.c_int64]
        self.lib, self.ref, self.visual = lib, ref, visual
        self.h = vp()
        assert lib.ref_gfx_create(C.byref(ref.cfg), C.byref(visual), C.byref(self.h)) == 0

    def close(self):
        if self.h:
            self.lib.ref_gfx_destroy(self.h)
            self.h = None

    def update(self, d_t=None):
        """-> (pixels (H, W) uint32, polylines (n_lines, length, 2) int32, segments (n, 4) int32)."""
        assert self.lib.ref_gfx_update(self.h, self.ref.h, self.ref.cfg.d_t if d_t is None else d_t) == 0
        n = self.lib.ref_gfx_pixels(None, 0)
        pixels = np.zeros(n, np.uint32)
        self.lib.ref_gfx_pixels(pixels.ctypes.data, n)
        nl = C.c_int32()
        n = self.lib.ref_gfx_polylines(None, 0, C.byref(nl))
        pts = np.zeros(n, np.int32)
        self.lib.ref_gfx_polylines(pts.ctypes.data, n, C.byref(nl))
        n = self.lib.ref_gfx_segments(None, 0)
        seg = np.zeros(n, np.int32)
        self.lib.ref_gfx_segments(seg.ctypes.data, n)
        H, W = self.ref.shape
        return (pixels.reshape(H, W), pts.reshape(nl.value, -1, 2) if nl.value else pts.reshape(0, 0, 2),
                seg.reshape(-1, 4))


def ref_sample_velocity(ref: RefSim, xs, ys):
    """Fluid::get_general_velocity evaluated by the reference's own device code."""
    lib = ref.lib
    vp = C.c_void_p
    lib.ref_sample_velocity.restype, lib.ref_sample_velocity.argtypes = C.c_int, [vp, C.c_int, vp, vp, vp, vp]
    xs = np.ascontiguousarray(xs, np.float32)
    ys = np.ascontiguousarray(ys, np.float32)
    ou, ov = np.empty_like(xs), np.empty_like(ys)
    assert lib.ref_sample_velocity(ref.h, xs.size, xs.ctypes.data, ys.ctypes.data, ou.ctypes.data, ov.ctypes.data) == 0
    return ou, ov
