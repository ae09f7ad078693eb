#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the OpenSayal step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

metric    cell-steps/s = (cells x steps) / device time, n = 50 SOR iterations per step (BASELINE.json)
workload  N = 1: BASELINE configs[1] — 1920x1080 wind tunnel (speed 200, pipe_height 270), disc r=36, smoke
          on, drain on, n=50, viscosity 0, synthetic initial fields.
          N > 1: the same slab per GPU stacked in y (1920 x 1080*N), one process per GPU, ghost-row exchange
          between neighbours ("weak" scaling).
value     whole-job throughput, state resident in HBM, every step timed on its own with CUDA events on the
          sim's stream and the L2 flushed (256 MiB write) between steps; max over ranks.
e2e       the same job through the public API from pinned HOST buffers: upload u, v, smoke, K x step with a
          host-side Source, download u, v, smoke — wall clock, copies inside the timed region.
roofline  the dominant kernel (projection_pack_kernel): algorithmic bytes per launch / mean launch duration
          (CUDA events around the projection stage), against MEASURED_PEAKS.json's HBM copy bandwidth.
cpu_baseline  the CPU restatement (oracle/, "port") on one host core, bounded sample of the same workload.

--impl reference: the reference's own implementation of the path.  OpenSayal has NO CPU path — its
implementation is CUDA (src/fluid.cu) — so this arm runs oracle/_ref (the unmodified reference sources built
for sm_100 with the reference's Release flags) on the same B200, same config, same timing protocol; if that
library is unavailable it falls back to the CPU port on all host threads and says so.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_SOR = 50


def algorithmic_bytes_per_cell_step(n, pressure, smoke, drag=False):
    """B_alg of SURVEY.md §8d / BASELINE.md §3."""
    return 8 + (8 if drag else 0) + n * (17 + 8 * pressure) + 17 + 17 * smoke


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.period = period
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv and not os.environ.get("SAYAL_BENCH_NO_SAMPLER"):
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_config(n_gpus: int, defaults=None):
    """defaults: see baseline_config — the reference arm builds its configuration without the product library."""
    from opensayal_b200.synthetic import baseline_config
    cfg = baseline_config(1, defaults=defaults)
    if n_gpus > 1:  # the same slab per GPU, stacked in y
        cfg = baseline_config(1, width=1920, height=1080 * n_gpus, defaults=defaults)
        cfg["sim.wind_tunnel.pipe_height"] = 270 * n_gpus
        cfg["sim.obstacle.radius"] = 36.0
    return cfg


def config_block(cfg, n_gpus, extra=None):
    c = cfg.c
    out = {"workload": f"{c.width}x{c.height} wind tunnel (speed {c.wt_speed:g}, pipe_height {c.wt_pipe_height}), "
                       f"disc r={c.obstacle_radius:g} at ({c.obstacle_center_x},{c.obstacle_center_y}), smoke on, drain on, "
                       f"n={c.proj_n}, o={c.proj_o:g}, d_t={c.d_t:g}, viscosity 0 — BASELINE configs[1]"
                       + (f" per GPU, stacked in y over {n_gpus} slabs" if n_gpus > 1 else ""),
           "grid": [c.width, c.height], "cells": c.width * c.height, "sor_iterations": c.proj_n,
           "parallelism": "single GPU" if n_gpus == 1 else f"y-slabs x{n_gpus}, ghost-row exchange",
           "l2": "flushed between timed steps (256 MiB write)",
           "initial_fields": "synthetic_fields(seed=1234): sin/cos modes amplitude 40 + noise 5"}
    if extra:
        out.update(extra)
    return out


# ----------------------------------------------------------------------------------------------------
def bench_ours_single(args):
    import numpy as np
    import torch

    from opensayal_b200 import Fluid, Source
    from opensayal_b200.synthetic import synthetic_fields

    torch.cuda.set_device(0)
    cfg = workload_config(1)
    c = cfg.c
    W, H = c.width, c.height
    cells = W * H
    u, v, sm = synthetic_fields(W, H)
    sim = Fluid(cfg, device=0)
    for name, a in (("u", u), ("v", v), ("smoke", sm)):
        sim.set_field(name, a)
    stream = torch.cuda.ExternalStream(sim.stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    sim.run(max(args.warmup, 3))
    sim.sync()

    # ---- device-resident throughput: every step timed on its own, L2 flushed in between
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = sim.launch_count
    with ClockSampler(0, period=0.004) as clocks:  # several samples inside a region of a few milliseconds
        done = 0
        while done < args.steps:  # the host enqueues a chunk behind the gate, then lets the device run it from its queue
            chunk = min(40, args.steps - done)
            gate = not os.environ.get("SAYAL_BENCH_NO_GATE")  # under ncu every launch is serialised: a gate would spin to its time-out
            if gate:
                sim.stream_hold()
                for _ in range(2):  # untimed: the first step behind the gate finds the GPU ramping up from the one-thread
                    flush_l2()      # spin (measured: step 0 took 0.22-0.27 ms, every other step 0.178) — same as bench_slabs
                    sim.run(1)
            for k in range(done, done + chunk):
                flush_l2()
                starts[k].record(stream)
                sim.run(1)
                stops[k].record(stream)
            if gate:
                sim.stream_release()
            sim.sync()
            done += chunk
    chunks = -(-args.steps // 40)
    extra = 0 if os.environ.get("SAYAL_BENCH_NO_GATE") else 2 * chunks  # untimed alignment steps behind each gate
    launches = (sim.launch_count - launches0) * args.steps // (args.steps + extra)  # launches of the timed steps only
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    ms_per_step = sum(step_ms) / len(step_ms)
    value = cells / (ms_per_step * 1e-3)

    # ---- the same loop without flushing (steady-state simulation: state stays in L2), for information
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.run(args.steps)
    e1.record(stream)
    sim.sync()
    warm_ms = e0.elapsed_time(e1) / args.steps

    # ---- dominant kernel: the projection stage alone (passes x projection_pack_kernel)
    n = c.proj_n
    l0 = sim.launch_count
    sim.stage_projection(n, c.d_t)
    sim.sync()
    passes = sim.launch_count - l0
    pstarts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pstops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        flush_l2()
        pstarts[k].record(stream)
        sim.stage_projection(n, c.d_t)
        pstops[k].record(stream)
    sim.sync()
    proj_ms = sum(s.elapsed_time(e) for s, e in zip(pstarts, pstops)) / args.steps
    launch_ms = proj_ms / passes
    alg_bytes_launch = 17.0 * cells * n / passes  # 17 B per cell per iteration x iterations per launch
    achieved = alg_bytes_launch / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    b_alg = algorithmic_bytes_per_cell_step(n, int(bool(c.enable_pressure)), int(bool(c.enable_smoke and c.wt_smoke != 0)))
    # physical DRAM bytes per launch of that kernel, from the committed ncu --set full capture (profiles/traffic.json,
    # written by tools/ncu_traffic.py); says whether the captured plan is the plan that just ran
    plan = {"temporal_block": sim.get_option("plan_temporal_block"), "tile_rows_per_warp": sim.get_option("plan_rows_per_warp"),
            "projection_kernel": sim.get_option("projection_kernel")}
    traffic, traffic_src, traffic_match, dram_frac = None, None, None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            t = json.loads(tp.read_text())
            if t.get("grid", [1920, 1080]) == [W, H]:
                traffic = t["dram_bytes_per_launch"]
                traffic_match = (t.get("temporal_block") == plan["temporal_block"]
                                 and t.get("rows_per_warp") == plan["tile_rows_per_warp"])
                traffic_src = (f"profiles/traffic.json: {t['source']}; captured with T={t['temporal_block']} "
                               f"rows/warp={t['rows_per_warp']}")
                # the physical figure beside the algorithmic one: ncu DRAM bytes / this run's launch time / peak
                dram_frac = round(traffic / (launch_ms * 1e-3) / 1e9 / peak, 4)
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "projection_pack_kernel", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "traffic_plan_matches": traffic_match, "dram_frac": dram_frac,
                "peak_source": peak_src, "launches_per_step": passes, "ms_per_launch": round(launch_ms, 5),
                "algorithmic_bytes_per_launch": alg_bytes_launch,
                "share_of_step": round(proj_ms / ms_per_step, 3),
                "whole_step": {"algorithmic_bytes_per_cell_step": b_alg,
                               "achieved_GBps": round(value * b_alg / 1e9, 1),
                               "frac": round(value * b_alg / 1e9 / peak, 4)},
                "note": "algorithmic bytes give no credit for temporal blocking (T iterations per HBM pass) or L2 "
                        "residency, so frac > 1 means fewer physical bytes than one pass per sweep; dram_frac is the "
                        "physical DRAM traffic of the same launch against the same peak; see DESIGN.md"}

    # ---- end to end from pinned host buffers through the public API: one batched upload, K steps with a host-side
    # Source each (main.cu:97), one batched download — wall clock, copies inside the timed region
    pinned = {k: torch.empty((H, W), dtype=torch.float32).pin_memory() for k in ("u", "v", "smoke")}
    for k, a in (("u", u), ("v", v), ("smoke", sm)):
        pinned[k].numpy()[...] = a
    out = {k: torch.empty((H, W), dtype=torch.float32).pin_memory() for k in ("u", "v", "smoke")}
    src = Source()  # inactive, passed from the host every step like main.cu:97 does
    field_bytes = 3 * cells * 4

    def e2e_once():
        sim.sync()
        t0 = time.perf_counter()
        sim.set_fields_from({k: t_.data_ptr() for k, t_ in pinned.items()})
        for _ in range(args.steps):
            sim.step_async(src, c.d_t)
        sim.get_fields_into({k: t_.data_ptr() for k, t_ in out.items()})
        return time.perf_counter() - t0

    e2e_once()  # untimed: first touch of the pinned pages by the copy engine
    e2e_s = min(e2e_once() for _ in range(3))
    assert np.isfinite(out["u"].numpy()).all()
    # the copies alone (same calls, no steps): what the PCIe link of this box gives for these buffers
    sim.sync()
    t0 = time.perf_counter()
    sim.set_fields_from({k: t_.data_ptr() for k, t_ in pinned.items()})
    sim.sync()
    t1 = time.perf_counter()
    sim.get_fields_into({k: t_.data_ptr() for k, t_ in out.items()})
    t2 = time.perf_counter()
    e2e_value = cells * args.steps / e2e_s
    e2e = {"value": e2e_value, "unit": "cell-steps/s", "h2d_bytes_per_step": field_bytes / args.steps + 20,
           "d2h_bytes_per_step": field_bytes / args.steps, "seconds": e2e_s,
           "h2d_GBps": round(field_bytes / (t1 - t0) / 1e9, 2), "d2h_GBps": round(field_bytes / (t2 - t1) / 1e9, 2),
           "protocol": "sayal_set_fields(u,v,smoke) from pinned host + K x sayal_step(host Source) + sayal_get_fields(u,v,smoke), "
                       "wall clock, best of 3"}

    strong = None
    if not os.environ.get("SAYAL_BENCH_SKIP_STRONG"):
        plan_text = sim.plan_log()
        sim.close()
        del flush
        torch.cuda.empty_cache()
        from opensayal_b200.slab import strong_16384_record
        strong = strong_16384_record(1, 0, 0)
    else:
        plan_text = sim.plan_log()
        sim.close()

    cpu = None if args.skip_cpu_baseline else cpu_baseline_port(cfg, threads=1, target_seconds=12.0)
    line = {
        "metric": "cell-steps/sec (n=50 SOR)", "value": value, "unit": "cell-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(cfg, 1),
        "plan": dict(plan, candidates=plan_text.strip().split("\n")),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks.summary(),
        "steady_state_ms_per_step_l2_warm": warm_ms,
        "step_ms": {"min": round(min(step_ms), 4), "median": round(statistics.median(step_ms), 4), "max": round(max(step_ms), 4),
                    "slowest_step": step_ms.index(max(step_ms))},
        "stage_ms": {"projection": proj_ms},
        "strong_16384": strong,
    }
    return line


def cpu_baseline_port(cfg, threads, target_seconds):
    """The oracle ("port": CPU restatement of fluid.cu) on `threads` host cores, bounded sample."""
    from opensayal_b200.synthetic import synthetic_fields
    from oracle.oracle import OracleSim
    c = cfg.c
    u, v, sm = synthetic_fields(c.width, c.height)
    o = OracleSim(c, threads=threads)
    for name, a in (("u", u), ("v", v), ("smoke", sm)):
        o.set_field(name, a)
    t0 = time.perf_counter()
    steps = 0
    while True:
        o.step(None, c.d_t)
        steps += 1
        if time.perf_counter() - t0 > target_seconds or steps >= 64:
            break
    dt = time.perf_counter() - t0
    o.close()
    return {"value": c.width * c.height * steps / dt, "unit": "cell-steps/s", "cores": threads, "kind": "port",
            "sample": f"{steps} full steps of the same {c.width}x{c.height} n={c.proj_n} workload, {dt:.1f} s, "
                      f"host has {os.cpu_count()} cores"}


# ----------------------------------------------------------------------------------------------------
def bench_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return None
    import torch
    from oracle.oracle import REF_LIB, RefSim
    from opensayal_b200.synthetic import synthetic_fields

    from oracle.oracle import reference_defaults
    # the reference is single-GPU (no multi-GPU code exists in it); its configuration is built without the product
    # library (reference_defaults), so this arm maps oracle/_ref only
    cfg = workload_config(1, defaults=reference_defaults)
    c = cfg.c
    cells = c.width * c.height
    warm = max(args.warmup, 3)
    if REF_LIB.exists() and torch.cuda.is_available():
        torch.cuda.set_device(0)
        u, v, sm = synthetic_fields(c.width, c.height)
        ref = RefSim(c, device=0)
        for name, a in (("u", u), ("v", v), ("smoke", sm)):
            ref.set_field(name, a)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ref.run_timed(warm)
        with ClockSampler(0) as clocks:
            ms = []
            for _ in range(args.steps):
                flush.fill_(1)
                torch.cuda.synchronize()
                ms.append(ref.run_timed(1))
        ms_per_step = sum(ms) / len(ms)
        value = cells / (ms_per_step * 1e-3)
        launches = 1 + 2 * c.proj_n + 1 + 2 + 3  # forces, 2n sweeps, extrapolation, 2 + 3 advection
        ref.close()
        # the same with the reference's shipped default fluid.viscosity = 0.001 (config_parser.cpp:117): n more
        # (racy) diffusion launches per step, fluid.cu:775-777 — reported beside the headline, not instead of it
        cfg_v = workload_config(1, defaults=reference_defaults)
        cfg_v["fluid.viscosity"] = 0.001
        ref = RefSim(cfg_v.c, device=0)
        for name, a in (("u", u), ("v", v), ("smoke", sm)):
            ref.set_field(name, a)
        ref.run_timed(warm)
        ms_v = []
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            ms_v.append(ref.run_timed(1))
        ms_visc = sum(ms_v) / len(ms_v)
        line = {
            "impl": "reference", "metric": "cell-steps/sec (n=50 SOR)", "value": value, "unit": "cell-steps/s",
            "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(cfg, 1),
            "reference_build": "oracle/_ref: /root/reference/src/fluid.cu + helper.cu unmodified, -O3 --use_fast_math "
                               "-rdc=true -gencode arch=compute_100,code=sm_100, block (64,1), fluid.viscosity=0 (H1)",
            "default_viscosity_0.001": {"ms_per_step": ms_visc, "value": cells / (ms_visc * 1e-3),
                                        "note": "same workload with the reference's shipped default fluid.viscosity "
                                                "(config_parser.cpp:117): + n diffusion launches per step"},
            "cpu_baseline": {"value": value, "unit": "cell-steps/s", "cores": 0, "kind": "reference",
                             "sample": f"{args.steps} full steps; OpenSayal has no CPU path, its implementation of the "
                                       "step is CUDA and ran on the same B200 (0 host cores on the data path)"},
            "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": launches * args.steps, "clocks": clocks.summary(),
        }
        ref.close()
        return line
    # fallback: CPU port with every host thread
    from oracle.oracle import oracle_lib
    threads = oracle_lib().oracle_max_threads()
    cpu = cpu_baseline_port(cfg, threads=threads, target_seconds=20.0)
    return {"impl": "reference", "metric": "cell-steps/sec (n=50 SOR)", "value": cpu["value"], "unit": "cell-steps/s",
            "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": cells / cpu["value"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(cfg, 1), "note": "oracle/_ref unavailable: CPU port on all host threads",
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only")
    args = ap.parse_args()
    if args.impl == "reference":
        line = bench_reference(args)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0
    rank, world, local = dist_env()
    if args.gpus > 1 or world > 1:
        from opensayal_b200.slab import bench_slabs
        line = bench_slabs(args, workload_config, config_block, ClockSampler, measured_hbm_peak,
                           algorithmic_bytes_per_cell_step)
    else:
        line = bench_ours_single(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
