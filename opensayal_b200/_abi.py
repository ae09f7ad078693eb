"""ctypes binding of include/sayal.h.

The shared library is the product: if it is missing, or was built without the CUDA kernels, importing a
symbol fails loudly — there is no Python or CPU fallback for the step path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libsayal_b200.so"

SAYAL_OK = 0
SAYAL_EINVAL, SAYAL_ECUDA, SAYAL_EIO, SAYAL_EPARSE, SAYAL_ENOMEM, SAYAL_EBUSY, SAYAL_ELINK = -1, -2, -3, -4, -5, -6, -7
LINK_INFO_BYTES = 256
U, V, P, SMOKE, IS_SOLID, TOTAL_S = range(6)
FIELD_NAMES = {"u": U, "v": V, "p": P, "smoke": SMOKE, "is_solid": IS_SOLID, "total_s": TOTAL_S}


class SayalConfig(C.Structure):
    """struct sayal_config — field order must match include/sayal.h."""

    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("cell_size", C.c_float),
        ("enable_drain", C.c_int32),
        ("enable_pressure", C.c_int32),
        ("enable_smoke", C.c_int32),
        ("enable_interactive", C.c_int32),
        ("proj_n", C.c_int32),
        ("proj_o", C.c_float),
        ("wt_pipe_height", C.c_int32),
        ("wt_pipe_length", C.c_int32),
        ("wt_smoke_length", C.c_int32),
        ("wt_smoke_height", C.c_int32),
        ("wt_smoke_count", C.c_int32),
        ("wt_speed", C.c_float),
        ("wt_smoke", C.c_float),
        ("g", C.c_float),
        ("d_t", C.c_float),
        ("enable_real_time", C.c_int32),
        ("real_time_multiplier", C.c_float),
        ("smoke_enable_decay", C.c_int32),
        ("smoke_decay_rate", C.c_float),
        ("obstacle_enable", C.c_int32),
        ("obstacle_center_x", C.c_int32),
        ("obstacle_center_y", C.c_int32),
        ("obstacle_radius", C.c_float),
        ("density", C.c_float),
        ("drag_coeff", C.c_float),
        ("viscosity", C.c_float),
        ("block_size_x", C.c_int32),
        ("block_size_y", C.c_int32),
    ]

    def copy(self) -> "SayalConfig":
        out = SayalConfig()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(SayalConfig))
        return out

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class SayalSource(C.Structure):
    _fields_ = [
        ("active", C.c_int32),
        ("smoke", C.c_float),
        ("velocity", C.c_float),
        ("x", C.c_int32),
        ("y", C.c_int32),
    ]


class SayalSlab(C.Structure):
    _fields_ = [
        ("global_height", C.c_int32),
        ("row0", C.c_int32),
        ("rows", C.c_int32),
        ("halo", C.c_int32),
    ]


class SayalVisual(C.Structure):
    """struct sayal_visual — the members of Config GraphicsHandler reads (graphics_handler.cu:99-121)."""

    _fields_ = [
        ("cell_pixel_size", C.c_int32),
        ("arrows_enable", C.c_int32),
        ("arrows_distance", C.c_int32),
        ("arrows_length_multiplier", C.c_float),
        ("arrows_disable_threshold", C.c_float),
        ("arrows_head_length", C.c_int32),
        ("path_line_enable", C.c_int32),
        ("path_line_length", C.c_int32),
        ("path_line_distance", C.c_int32),
        ("arrows_color", C.c_int32 * 4),
        ("path_line_color", C.c_int32 * 4),
    ]


class SayalArrow(C.Structure):
    """struct sayal_arrow == ArrowData (graphics_handler.cuh:17-27) with an int32 `valid`."""

    _fields_ = [(n, C.c_int32) for n in ("start_x", "start_y", "end_x", "end_y", "right_head_end_x", "right_head_end_y",
                                         "left_head_end_x", "left_head_end_y", "valid")]


ARROW_DTYPE = [(n, "<i4") for n, _ in SayalArrow._fields_]

# every symbol include/sayal.h declares: (name, restype, argtypes)
_cfgp = C.POINTER(SayalConfig)
_srcp = C.POINTER(SayalSource)
_simp = C.c_void_p
_visp = C.POINTER(SayalVisual)
_i32p = C.POINTER(C.c_int32)
SYMBOLS = [
    ("sayal_visual_defaults", C.c_int, [_visp]),
    ("sayal_visual_load", C.c_int, [C.c_char_p, _visp]),
    ("sayal_visual_parse", C.c_int, [C.c_char_p, C.c_size_t, _visp]),
    ("sayal_stage_diffusion", C.c_int, [_simp, C.c_int32, C.c_float]),
    ("sayal_render_pixels", C.c_int, [_simp, C.c_void_p]),
    ("sayal_frame_submit", C.c_int, [_simp]),
    ("sayal_frame_acquire", C.c_int, [_simp, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    ("sayal_arrows", C.c_int, [_simp, _visp, C.c_void_p, C.c_int32, _i32p, _i32p]),
    ("sayal_path_lines", C.c_int, [_simp, _visp, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, _i32p, _i32p]),
    ("sayal_config_defaults", C.c_int, [C.c_int32, C.c_int32, _cfgp]),
    ("sayal_config_load", C.c_int, [C.c_char_p, _cfgp]),
    ("sayal_config_parse", C.c_int, [C.c_char_p, C.c_size_t, _cfgp]),
    ("sayal_create", C.c_int, [_cfgp, C.c_int32, C.POINTER(_simp)]),
    ("sayal_create_slab", C.c_int, [_cfgp, C.c_int32, C.POINTER(SayalSlab), C.POINTER(_simp)]),
    ("sayal_destroy", None, [_simp]),
    ("sayal_step", C.c_int, [_simp, _srcp, C.c_float]),
    ("sayal_run", C.c_int, [_simp, C.c_int32, C.c_float]),
    ("sayal_sync", C.c_int, [_simp]),
    ("sayal_get_field", C.c_int, [_simp, C.c_int32, C.c_void_p]),
    ("sayal_set_field", C.c_int, [_simp, C.c_int32, C.c_void_p]),
    ("sayal_get_field_device", C.c_int, [_simp, C.c_int32, C.c_void_p]),
    ("sayal_set_field_device", C.c_int, [_simp, C.c_int32, C.c_void_p]),
    ("sayal_get_fields", C.c_int, [_simp, C.c_int32, _i32p, C.POINTER(C.c_void_p)]),
    ("sayal_set_fields", C.c_int, [_simp, C.c_int32, _i32p, C.POINTER(C.c_void_p)]),
    ("sayal_device_ptr", C.c_int, [_simp, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("sayal_pressure_range", C.c_int, [_simp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    ("sayal_sample_velocity", C.c_int, [_simp, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("sayal_stage_forces", C.c_int, [_simp, _srcp, C.c_float]),
    ("sayal_stage_zero_pressure", C.c_int, [_simp]),
    ("sayal_stage_projection", C.c_int, [_simp, C.c_int32, C.c_float]),
    ("sayal_stage_extrapolation", C.c_int, [_simp]),
    ("sayal_stage_advect_velocity", C.c_int, [_simp, C.c_float]),
    ("sayal_stage_advect_smoke", C.c_int, [_simp, C.c_float]),
    ("sayal_set_option", C.c_int, [_simp, C.c_char_p, C.c_int64]),
    ("sayal_get_option", C.c_int, [_simp, C.c_char_p, C.POINTER(C.c_int64)]),
    ("sayal_debug_timeline", C.c_int, [_simp, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    ("sayal_debug_link_words", C.c_int, [_simp, C.c_void_p]),
    ("sayal_debug_stage_times", C.c_int, [_simp, C.c_void_p, C.c_int32]),
    ("sayal_debug_tile_list", C.c_int, [_simp, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    ("sayal_stream_delay", C.c_int, [_simp, C.c_int64]),
    ("sayal_stream_hold", C.c_int, [_simp]),
    ("sayal_stream_release", C.c_int, [_simp]),
    ("sayal_plan_log", C.c_int, [_simp, C.c_char_p, C.c_int32]),
    ("sayal_debug_pass_plans", C.c_int, [C.c_int32] * 8 + [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    ("sayal_launch_count", C.c_int64, [_simp]),
    ("sayal_stream", C.c_void_p, [_simp]),
    ("sayal_slab_pack_edge", C.c_int, [_simp, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    ("sayal_slab_unpack_ghost", C.c_int, [_simp, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    ("sayal_slab_ipc_export", C.c_int, [_simp, C.c_void_p]),
    ("sayal_slab_ipc_connect", C.c_int, [_simp, C.c_int32, C.c_void_p]),
    ("sayal_slab_connect_local", C.c_int, [_simp, C.c_int32, _simp]),
    ("sayal_slab_exchange", C.c_int, [_simp, C.c_int32]),
    ("sayal_last_error", C.c_char_p, []),
    ("sayal_abi_version", C.c_int, []),
]

ABI_VERSION = 3
_lib = None


class SayalError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"sayal error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Load libsayal_b200.so (built by opensayal_b200/csrc/Makefile or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SAYAL_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: the CUDA extension is the product and has no fallback. "
            "Build it with `make -C opensayal_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(str(path))
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sayal_abi_version() != ABI_VERSION:
        raise ImportError("libsayal_b200.so: ABI version mismatch")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != SAYAL_OK:
        msg = load().sayal_last_error()
        raise SayalError(code, msg.decode() if msg else "")
