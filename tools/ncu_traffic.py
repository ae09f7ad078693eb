#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernel from an `ncu --set full` capture (read here, no GPU):
   python tools/ncu_traffic.py gpurun_out/<tag>_projection.ncu-rep gpurun_out/<tag>_plan.json profiles/traffic.json
Writes {"kernel", "temporal_block", "rows_per_warp", "dram_bytes_per_launch", "launches", "source"}; bench.py puts
dram_bytes_per_launch into roofline.traffic when the plan it runs matches."""
import csv, io, json, subprocess, sys

rep, plan_json, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: k for k, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
vals, times, names = [], [], set()
for r in data:
    if "projection_pack" not in r[col["Kernel Name"]]:
        continue
    names.add(r[col["Kernel Name"]].split("(")[0][-60:])
    b = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[k]].replace(",", "")) * scale[units[col[k]]]
    vals.append(b)
    times.append(float(r[col["gpu__time_duration.sum"]].replace(",", "")))
plan = json.load(open(plan_json))["plan"]
res = {"kernel": sorted(names)[0] if names else None, "temporal_block": plan["temporal_block"],
       "rows_per_warp": plan["tile_rows_per_warp"], "dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
       "ncu_time_us_per_launch": sum(times) / len(times), "source": rep.split("/")[-1] + " (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches)"}
json.dump(res, open(out, "w"), indent=1)
print(res)
