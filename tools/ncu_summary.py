#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/x.ncu-rep [--all]
Prints, per captured launch, the metrics the roofline / stall analysis in DESIGN.md cites."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum",
    "smsp__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]


def main():
    path = sys.argv[1]
    show_all = "--all" in sys.argv
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: k for k, h in enumerate(hdr)}
    for r in data:
        print("=" * 100)
        print(r[col["Kernel Name"]][:100])
        for k in KEYS:
            if k in col:
                print(f"  {k:75s} {r[col[k]]:>18s} {units[col[k]]}")
        stalls = [(float(r[c].replace(",", "") or 0), h) for h, c in col.items()
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue")
        if show_all:
            for h, c in col.items():
                print(f"    {h:90s} {r[c]:>18s} {units[c]}")


if __name__ == "__main__":
    main()
