#!/usr/bin/env python
"""Executed-instruction mix of one kernel in an .ncu-rep: python tools/ncu_instmix.py rep regex [--hot N]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
hot = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches may match: take the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
hdr = b["rows"][0]; col = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter(); total = 0; samples = collections.Counter(); tot_s = 0
lines = []
for r in b["rows"][1:]:
    if len(r) < len(hdr): continue
    op = r[col["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    name = op[0].split(".")[0] if op else "?"
    n = int(r[col["Instructions Executed"]] or 0)
    s = int(r[col["# Samples"]] or 0)
    mix[name] += n; total += n; samples[name] += s; tot_s += s
    lines.append((s, n, r[col["Source"]].strip()))
print(b["name"]); print("total warp-instructions", total, " samples", tot_s)
for k, v in mix.most_common(25):
    print(f"  {k:10s} {v:10d} {100*v/total:5.1f}%   samples {100*samples[k]/max(tot_s,1):5.1f}%")
if hot:
    print("hottest instructions by stall samples:")
    for s, n, src in sorted(lines, reverse=True)[:hot]:
        print(f"  {s:6d} {n:9d}  {src[:90]}")
