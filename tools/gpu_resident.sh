#!/bin/bash
# tools/gpu_resident.sh <tag> — resident projection: parity tests, then the bench with resident plans off / only / tuned
set -u
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "resident" 2>&1 | tail -15
for mode in 0 2 1; do
  SAYAL_RESIDENT=$mode SAYAL_BENCH_SKIP_STRONG=1 timeout 300 python bench.py --skip-cpu-baseline > gpurun_out/${tag}_bench_resident$mode.json 2> gpurun_out/${tag}_bench_resident$mode.err
  echo "resident=$mode rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_resident$mode.json"))
print("ms/step", round(d["ms_per_step"],4), "value", round(d["value"]/1e9,3), "proj ms", d["stage_ms"], "plan", d["plan"]["temporal_block"], d["plan"]["tile_rows_per_warp"], "e2e", round(d["e2e"]["value"]/1e9,3))
print("\n".join(d["plan"]["candidates"][:12]))
PY
done
