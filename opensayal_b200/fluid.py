"""Host-side mirror of the reference's interface for the step path.

Reference                                   here
---------------------------------------     ----------------------------------------------
ConfigParser(file).parse() -> Config        ConfigParser(file).parse() -> Config     (config_parser.cpp:11-185)
struct Source {active,smoke,velocity,pos}   Source(active, smoke, velocity, position) (fluid.cuh:8-13)
Fluid fluid(config)                         Fluid(config, device=0)                   (fluid.cu:41-71)
fluid.update(source, d_t)                   fluid.update(source, d_t)                 (fluid.cu:770-795)
fluid.width / height / min_pressure ...     same attribute names                      (fluid.cuh:40-62)
fluid.d_vel_x ... (device pointers)         fluid.device_ptr("u") / fluid.vel_x (host copy, H x W)
d_fluid->get_general_velocity(x, y)         fluid.get_general_velocity(xs, ys)        (fluid.cu:541-545)

Everything here is a thin ctypes layer over libsayal_b200.so; no arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import SayalConfig, SayalSlab, SayalSource, SayalVisual, check, load

# JSON key of the reference -> member of sayal_config
_KEYS = {
    "sim.width": "width", "sim.height": "height", "sim.cell_size": "cell_size",
    "sim.enable_drain": "enable_drain", "sim.enable_pressure": "enable_pressure",
    "sim.enable_smoke": "enable_smoke", "sim.enable_interactive": "enable_interactive",
    "sim.projection.n": "proj_n", "sim.projection.o": "proj_o",
    "sim.wind_tunnel.pipe_height": "wt_pipe_height", "sim.wind_tunnel.pipe_length": "wt_pipe_length",
    "sim.wind_tunnel.smoke_length": "wt_smoke_length", "sim.wind_tunnel.smoke_height": "wt_smoke_height",
    "sim.wind_tunnel.smoke_count": "wt_smoke_count", "sim.wind_tunnel.speed": "wt_speed",
    "sim.wind_tunnel.smoke": "wt_smoke", "sim.physics.g": "g", "sim.time.d_t": "d_t",
    "sim.time.enable_real_time": "enable_real_time", "sim.time.real_time_multiplier": "real_time_multiplier",
    "sim.smoke.enable_decay": "smoke_enable_decay", "sim.smoke.decay_rate": "smoke_decay_rate",
    "sim.obstacle.enable": "obstacle_enable", "sim.obstacle.center_x": "obstacle_center_x",
    "sim.obstacle.center_y": "obstacle_center_y", "sim.obstacle.radius": "obstacle_radius",
    "fluid.density": "density", "fluid.drag_coeff": "drag_coeff", "fluid.viscosity": "viscosity",
    "thread.cuda.block_size_x": "block_size_x", "thread.cuda.block_size_y": "block_size_y",
}


class _View:
    """config.sim.projection.n style access onto the flat C struct."""

    def __init__(self, cfg: "Config", prefix: str):
        object.__setattr__(self, "_cfg", cfg)
        object.__setattr__(self, "_prefix", prefix)

    def __getattr__(self, name):
        key = f"{self._prefix}.{name}"
        if key in _KEYS:
            return getattr(self._cfg.c, _KEYS[key])
        if any(k.startswith(key + ".") for k in _KEYS):
            return _View(self._cfg, key)
        raise AttributeError(key)

    def __setattr__(self, name, value):
        key = f"{self._prefix}.{name}"
        if key not in _KEYS:
            raise AttributeError(key)
        setattr(self._cfg.c, _KEYS[key], value)


class Config:
    """The simulation subset of the reference's `Config` tree (config_parser.hpp:9-114)."""

    def __init__(self, c: Optional[SayalConfig] = None, width: int = 1920, height: int = 1080):
        if c is None:
            c = SayalConfig()
            check(load().sayal_config_defaults(width, height, C.byref(c)))
        self.c = c

    @classmethod
    def defaults(cls, width: int = 1920, height: int = 1080, _struct=None, **overrides) -> "Config":
        """_struct: a ready-made SayalConfig to start from instead of asking the library for the defaults."""
        cfg = cls(_struct(width, height), width, height) if _struct is not None else cls(width=width, height=height)
        for key, value in overrides.items():
            cfg[key] = value
        return cfg

    def __getitem__(self, key: str):
        return getattr(self.c, _KEYS.get(key, key))

    def __setitem__(self, key: str, value):
        name = _KEYS.get(key, key)
        if not hasattr(self.c, name):
            raise KeyError(key)
        setattr(self.c, name, value)

    @property
    def sim(self):
        return _View(self, "sim")

    @property
    def fluid(self):
        return _View(self, "fluid")

    @property
    def thread(self):
        return _View(self, "thread")

    def copy(self) -> "Config":
        return Config(self.c.copy())


class ConfigParser:
    """ConfigParser (config_parser.hpp:116-129): reads ./OpenSayal.conf.json unless told otherwise."""

    def __init__(self, config_file_name: str = "OpenSayal.conf.json"):
        self.config_file_name = config_file_name

    def parse(self) -> Config:
        c = SayalConfig()
        check(load().sayal_config_load(str(self.config_file_name).encode(), C.byref(c)))
        return Config(c)

    @staticmethod
    def parse_text(text: str) -> Config:
        c = SayalConfig()
        raw = text.encode()
        check(load().sayal_config_parse(raw, len(raw), C.byref(c)))
        return Config(c)


class Visual:
    """The members of `Config` GraphicsHandler reads (graphics_handler.cu:99-121): sim.cell_pixel_size,
    visual.arrows.*, visual.path_line.*; defaults of config_parser.cpp:40, 120-181."""

    def __init__(self, v: Optional[SayalVisual] = None, **overrides):
        if v is None:
            v = SayalVisual()
            check(load().sayal_visual_defaults(C.byref(v)))
        self.v = v
        for key, value in overrides.items():
            setattr(self.v, key, value)

    @staticmethod
    def parse_text(text: str) -> "Visual":
        v = SayalVisual()
        raw = text.encode()
        check(load().sayal_visual_parse(raw, len(raw), C.byref(v)))
        return Visual(v)

    @staticmethod
    def parse_file(path: str) -> "Visual":
        v = SayalVisual()
        check(load().sayal_visual_load(str(path).encode(), C.byref(v)))
        return Visual(v)


@dataclass
class Source:
    """struct Source (fluid.cuh:8-13); default = the inactive source a headless run passes."""

    active: bool = False
    smoke: float = 0.0
    velocity: float = 0.0
    position: Tuple[int, int] = (0, 0)

    def _c(self) -> SayalSource:
        return SayalSource(int(bool(self.active)), float(self.smoke), float(self.velocity),
                           int(self.position[0]), int(self.position[1]))


class Fluid:
    """class Fluid (fluid.cuh:15-116) over the B200-native library."""

    def __init__(self, config: Config, device: int = 0, slab: Optional[Tuple[int, int, int]] = None):
        self._lib = load()
        self._sim = C.c_void_p()
        self.config = config.copy()
        c = self.config.c
        if slab is None:
            check(self._lib.sayal_create(C.byref(c), device, C.byref(self._sim)))
            self.rows = c.height
        else:
            row0, rows, halo = slab
            s = SayalSlab(c.height, row0, rows, halo)
            check(self._lib.sayal_create_slab(C.byref(c), device, C.byref(s), C.byref(self._sim)))
            self.rows = rows
        self.device = device
        # public const members of the reference class
        self.width, self.height = c.width, c.height
        self.g, self.density, self.viscosity, self.o = c.g, c.density, c.viscosity, c.proj_o
        self.cell_size, self.n, self.drag_coeff = int(c.cell_size), c.proj_n, c.drag_coeff
        self.enable_pressure, self.enable_smoke = bool(c.enable_pressure), bool(c.enable_smoke)
        self.enable_smoke_decay, self.smoke_decay_rate = bool(c.smoke_enable_decay), c.smoke_decay_rate

    # ---- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_sim", None) is not None and self._sim.value:
            self._lib.sayal_destroy(self._sim)
            self._sim = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- stepping ------------------------------------------------------------------------------
    def update(self, source: Optional[Source] = None, d_t: Optional[float] = None) -> None:
        """Fluid::update(Source, d_t): one step, returning when the device has finished (fluid.cu:794)."""
        self.step_async(source, d_t)
        self.sync()

    def step_async(self, source: Optional[Source] = None, d_t: Optional[float] = None) -> None:
        d_t = self.config.c.d_t if d_t is None else d_t
        src = source._c() if source is not None else None
        check(self._lib.sayal_step(self._sim, C.byref(src) if src is not None else None, d_t))

    def run(self, steps: int, d_t: Optional[float] = None) -> None:
        """Headless batch of `steps` updates with an inactive source; asynchronous."""
        check(self._lib.sayal_run(self._sim, steps, self.config.c.d_t if d_t is None else d_t))

    def sync(self) -> None:
        check(self._lib.sayal_sync(self._sim))

    # ---- staged access -------------------------------------------------------------------------
    def stage_forces(self, source: Optional[Source], d_t: float):
        src = source._c() if source is not None else None
        check(self._lib.sayal_stage_forces(self._sim, C.byref(src) if src is not None else None, d_t))

    def stage_zero_pressure(self):
        check(self._lib.sayal_stage_zero_pressure(self._sim))

    def stage_projection(self, iterations: int, d_t: float):
        check(self._lib.sayal_stage_projection(self._sim, iterations, d_t))

    def stage_diffusion(self, iterations: int, d_t: float):
        check(self._lib.sayal_stage_diffusion(self._sim, iterations, d_t))

    def stage_extrapolation(self):
        check(self._lib.sayal_stage_extrapolation(self._sim))

    def stage_advect_velocity(self, d_t: float):
        check(self._lib.sayal_stage_advect_velocity(self._sim, d_t))

    def stage_advect_smoke(self, d_t: float):
        check(self._lib.sayal_stage_advect_smoke(self._sim, d_t))

    # ---- state ---------------------------------------------------------------------------------
    def get_field(self, name: str) -> np.ndarray:
        """Host copy, shape (rows, W), reference layout: row 0 is j = H-1 (fluid.cu:163-165)."""
        fid = _abi.FIELD_NAMES[name]
        dtype = np.int32 if fid in (_abi.IS_SOLID, _abi.TOTAL_S) else np.float32
        out = np.empty((self.rows, self.width), dtype=dtype)
        check(self._lib.sayal_get_field(self._sim, fid, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_field(self, name: str, array: np.ndarray) -> None:
        fid = _abi.FIELD_NAMES[name]
        a = np.ascontiguousarray(array, dtype=np.float32)
        if a.shape != (self.rows, self.width):
            raise ValueError(f"{name}: expected shape {(self.rows, self.width)}, got {a.shape}")
        check(self._lib.sayal_set_field(self._sim, fid, a.ctypes.data_as(C.c_void_p)))

    vel_x = property(lambda self: self.get_field("u"))
    vel_y = property(lambda self: self.get_field("v"))
    pressure = property(lambda self: self.get_field("p"))
    smoke = property(lambda self: self.get_field("smoke"))
    is_solid = property(lambda self: self.get_field("is_solid"))
    total_s = property(lambda self: self.get_field("total_s"))

    def device_ptr(self, name: str):
        """(device pointer, pitch in elements, first global memory row, rows held)."""
        p, pitch, first, rows = C.c_void_p(), C.c_int64(), C.c_int32(), C.c_int32()
        check(self._lib.sayal_device_ptr(self._sim, _abi.FIELD_NAMES[name], C.byref(p), C.byref(pitch),
                                         C.byref(first), C.byref(rows)))
        return p.value, pitch.value, first.value, rows.value

    def _range(self):
        mn, mx = C.c_float(), C.c_float()
        check(self._lib.sayal_pressure_range(self._sim, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    min_pressure = property(lambda self: self._range()[0])
    max_pressure = property(lambda self: self._range()[1])

    def get_general_velocity(self, xs: Sequence[float], ys: Sequence[float]):
        xs = np.ascontiguousarray(xs, dtype=np.float32)
        ys = np.ascontiguousarray(ys, dtype=np.float32)
        ou, ov = np.empty_like(xs), np.empty_like(ys)
        check(self._lib.sayal_sample_velocity(self._sim, xs.size, xs.ctypes.data_as(C.c_void_p),
                                              ys.ctypes.data_as(C.c_void_p), ou.ctypes.data_as(C.c_void_p),
                                              ov.ctypes.data_as(C.c_void_p)))
        return ou, ov

    # ---- what GraphicsHandler::update computes from the fluid (graphics_handler.cu:463-478) ------------
    def render_pixels(self, prefill: int = 0) -> np.ndarray:
        """update_fluid_pixels: (rows, W) uint32 RGBA8888 frame of the current state, synchronous."""
        out = np.full((self.rows, self.width), prefill, dtype=np.uint32)
        check(self._lib.sayal_render_pixels(self._sim, out.ctypes.data_as(C.c_void_p)))
        return out

    def frame_submit(self) -> None:
        """Render the state as of the work enqueued so far and start its copy to pinned host memory; returns at once."""
        check(self._lib.sayal_frame_submit(self._sim))

    def frame_acquire(self, copy: bool = True):
        """(frame, step index) of the oldest submitted frame; waits for that copy only."""
        p, step = C.c_void_p(), C.c_int64()
        check(self._lib.sayal_frame_acquire(self._sim, C.byref(p), C.byref(step)))
        buf = (C.c_uint32 * (self.rows * self.width)).from_address(p.value)
        a = np.ctypeslib.as_array(buf).reshape(self.rows, self.width)
        return (a.copy() if copy else a), step.value

    def arrows(self, visual: "Visual") -> np.ndarray:
        """update_center_velocity_arrow: structured array (n_y, n_x) of ArrowData, row 0 = top of the picture."""
        nx, ny = C.c_int32(), C.c_int32()
        check(self._lib.sayal_arrows(self._sim, C.byref(visual.v), None, 0, C.byref(nx), C.byref(ny)))
        out = np.zeros((ny.value, nx.value), dtype=_abi.ARROW_DTYPE)
        if out.size:
            check(self._lib.sayal_arrows(self._sim, C.byref(visual.v), out.ctypes.data_as(C.c_void_p), out.size,
                                         C.byref(nx), C.byref(ny)))
        return out

    def path_lines(self, visual: "Visual", d_t: Optional[float] = None):
        """update_traces: (xs, ys) int32 arrays of shape (n_y, n_x, length); lines from solid cells are -1."""
        d_t = self.config.c.d_t if d_t is None else d_t
        nx, ny = C.c_int32(), C.c_int32()
        check(self._lib.sayal_path_lines(self._sim, C.byref(visual.v), d_t, None, None, 0, C.byref(nx), C.byref(ny)))
        shape = (ny.value, nx.value, visual.v.path_line_length)
        xs, ys = np.zeros(shape, np.int32), np.zeros(shape, np.int32)
        if xs.size:
            check(self._lib.sayal_path_lines(self._sim, C.byref(visual.v), d_t, xs.ctypes.data_as(C.c_void_p),
                                             ys.ctypes.data_as(C.c_void_p), xs.size, C.byref(nx), C.byref(ny)))
        return xs, ys

    # ---- tuning / introspection ----------------------------------------------------------------
    def set_option(self, key: str, value: int) -> None:
        check(self._lib.sayal_set_option(self._sim, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        check(self._lib.sayal_get_option(self._sim, key.encode(), C.byref(v)))
        return v.value

    def debug_timeline(self, max_tiles: int = 65536) -> np.ndarray:
        """Profiling only: (tiles, 5) int64 {entry, loaded, swept, stored (ns), SM id} of the last projection pass."""
        out = np.zeros((max_tiles, 5), dtype=np.int64)
        n = C.c_int32()
        check(self._lib.sayal_debug_timeline(self._sim, out.ctypes.data_as(C.c_void_p), max_tiles, C.byref(n)))
        return out[: n.value]

    def debug_tile_list(self, iterations_per_pass: int, capacity: int = 4096) -> np.ndarray:
        """Diagnostics: (tiles, 5) int32 {X0, Y0, y_end, vy0, vy1} of the plan's explicit tile list (may be empty)."""
        out = np.zeros((capacity, 5), dtype=np.int32)
        n = C.c_int32()
        check(self._lib.sayal_debug_tile_list(self._sim, int(iterations_per_pass), out.ctypes.data_as(C.c_void_p), capacity, C.byref(n)))
        return out[: n.value]

    def debug_stage_times(self) -> np.ndarray:
        """Profiling only (option debug_events): ms from the start of the last eager step to its stage boundaries."""
        out = np.zeros(10, dtype=np.float32)
        check(self._lib.sayal_debug_stage_times(self._sim, out.ctypes.data_as(C.c_void_p), 10))
        return out

    def debug_link_words(self) -> np.ndarray:
        """Diagnostics: 64 control words of the slab link + 32 private counters (csrc/sayal_internal.h)."""
        out = np.zeros(96, dtype=np.uint32)
        check(self._lib.sayal_debug_link_words(self._sim, out.ctypes.data_as(C.c_void_p)))
        return out

    def stream_delay(self, microseconds: int) -> None:
        """Measurement aid: hold the sim's stream so that a timed region can be enqueued ahead of the device."""
        check(self._lib.sayal_stream_delay(self._sim, int(microseconds)))

    def stream_hold(self) -> None:
        """Everything enqueued after this call waits on the device until stream_release()."""
        check(self._lib.sayal_stream_hold(self._sim))

    def stream_release(self) -> None:
        check(self._lib.sayal_stream_release(self._sim))

    def plan_log(self) -> str:
        """The tile plans the tuner considered for the last projection it planned (marked: the chosen one)."""
        buf = C.create_string_buffer(4096)
        check(min(self._lib.sayal_plan_log(self._sim, buf, 4096), 0))
        return buf.value.decode()

    def set_field_device(self, name: str, dev_ptr: int) -> None:
        """Owned rows from a device buffer in the reference layout (row pitch W); asynchronous on the sim's stream."""
        check(self._lib.sayal_set_field_device(self._sim, _abi.FIELD_NAMES[name], C.c_void_p(dev_ptr)))

    def get_field_device(self, name: str, dev_ptr: int) -> None:
        check(self._lib.sayal_get_field_device(self._sim, _abi.FIELD_NAMES[name], C.c_void_p(dev_ptr)))

    def set_fields_from(self, pointers: dict) -> None:
        """sayal_set_fields: {name: host address} — asynchronous for pinned memory (keep the buffers until sync)."""
        n = len(pointers)
        ids = (C.c_int32 * n)(*[_abi.FIELD_NAMES[k] for k in pointers])
        ptrs = (C.c_void_p * n)(*[int(v) for v in pointers.values()])
        check(self._lib.sayal_set_fields(self._sim, n, ids, ptrs))

    def get_fields_into(self, pointers: dict) -> None:
        """sayal_get_fields: {name: host address}; returns when every copy has landed."""
        n = len(pointers)
        ids = (C.c_int32 * n)(*[_abi.FIELD_NAMES[k] for k in pointers])
        ptrs = (C.c_void_p * n)(*[int(v) for v in pointers.values()])
        check(self._lib.sayal_get_fields(self._sim, n, ids, ptrs))

    @property
    def launch_count(self) -> int:
        return self._lib.sayal_launch_count(self._sim)

    @property
    def stream(self) -> int:
        return self._lib.sayal_stream(self._sim) or 0

    # slab links (NVLink peer memory; see slab.py and csrc/slab_exchange.cu)
    def ipc_export(self) -> bytes:
        """The opaque blob a neighbouring rank needs to reach this slab (IPC handles + geometry)."""
        buf = C.create_string_buffer(_abi.LINK_INFO_BYTES)
        check(self._lib.sayal_slab_ipc_export(self._sim, buf))
        return buf.raw

    def ipc_connect(self, side: int, info: bytes) -> None:
        buf = C.create_string_buffer(info, _abi.LINK_INFO_BYTES)
        check(self._lib.sayal_slab_ipc_connect(self._sim, side, buf))

    def connect_local(self, side: int, neighbour: "Fluid") -> None:
        check(self._lib.sayal_slab_connect_local(self._sim, side, neighbour._sim))

    def slab_exchange(self, field_mask: int) -> None:
        check(self._lib.sayal_slab_exchange(self._sim, field_mask))

    # slab plumbing (see slab.py)
    def pack_edge(self, side: int, nrows: int, field_mask: int, dev_ptr: int):
        check(self._lib.sayal_slab_pack_edge(self._sim, side, nrows, field_mask, C.c_void_p(dev_ptr)))

    def unpack_ghost(self, side: int, nrows: int, field_mask: int, dev_ptr: int):
        check(self._lib.sayal_slab_unpack_ghost(self._sim, side, nrows, field_mask, C.c_void_p(dev_ptr)))
