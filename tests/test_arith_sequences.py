"""CPU check of the instruction sequences the smoke advection issues in place of __fdiv_rn / __fsqrt_rn / __drcp_rn
(opensayal_b200/csrc/advect_tile.cu): tests/arith_sequences.c restates them with fmaf / fma and compares with the host's
IEEE operations over random operands in the kernel's guarded ranges.  No GPU, no product code."""
import shutil
import subprocess
from pathlib import Path

import pytest

SRC = Path(__file__).with_name("arith_sequences.c")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_sequences_round_correctly(tmp_path):
    exe = tmp_path / "arith_sequences"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(SRC), "-lm"], check=True)
    out = subprocess.run([str(exe), "1000000"], capture_output=True, text=True, timeout=120)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
    lines = out.stdout.splitlines()
    assert len(lines) == 15
    # divide and FP64 reciprocal: no mismatch anywhere in the band; square root: none from the exact first guess
    for line in lines:
        if line.startswith(("div", "drcp")) or "+0 ulp" in line:
            assert " 0 mismatches" in line, line
