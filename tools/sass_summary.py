#!/usr/bin/env python
"""SASS evidence for the kernels of libsayal_b200.so (no GPU needed): per kernel, the number of instructions and the
count of the mnemonics that matter for the design claims — packed fp32 (FADD2 / FMUL2 / FFMA2), shuffles, barriers,
shared / global / local memory accesses, FP64, TMA (UTMALDG / UBLKCP: none — the tiles go global -> registers, DESIGN.md
§4.1), PDL (ACQBULK) — plus registers and spill bytes from the ptxas logs of the build.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "opensayal_b200" / "libsayal_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
arch = set()
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        counts[kern][m.group(1)] += 1
        counts[kern]["_all"] += 1

demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(counts, demangle))
regs = {}
for log in (ROOT / "opensayal_b200" / "csrc" / "build").glob("*.ptxas.log"):
    cur = None
    for line in log.read_text().splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            regs.setdefault(cur, {})["spill"] = (int(m.group(1)), int(m.group(2)))
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            regs.setdefault(cur, {})["regs"] = int(m.group(1))

cols = ["FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA", "SHFL", "BAR", "LDS", "STS", "LDG", "STG", "LDL", "STL", "DFMA", "MUFU",
        "UTMALDG", "UBLKCP", "ACQBULK"]
print(f"# {lib.name}: cubins for {sorted(arch)}; columns: instructions, registers, spill bytes (stores/loads), then mnemonic counts")
print("kernel".ljust(86) + " instr regs spill      " + " ".join(c.rjust(6) for c in cols))
for k, c in counts.items():
    n = names.get(k, k)
    n = re.sub(r"sayal::\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    r = regs.get(k, {})
    sp = r.get("spill", (0, 0))
    print(n[:85].ljust(86) + f"{c['_all']:6d} {r.get('regs', 0):4d} {sp[0]:4d}/{sp[1]:<5d}" + " ".join(str(c[m]).rjust(6) for m in cols))
