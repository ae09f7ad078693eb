#!/bin/bash
# scratch: GPU tests + stage breakdown + two bench runs
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_slab_gpu.py tests/test_visual.py -m gpu -x -q 2>&1 | tail -2
python tools/step_breakdown.py 2>&1 | tail -2
for m in 1 2; do
  SAYAL_BENCH_SKIP_STRONG=1 python bench.py --skip-cpu-baseline --steps 40 > gpurun_out/tmp_bench_$m.json
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_bench_$m.json"))
print("ms/step", round(d["ms_per_step"],4), "warm", round(d["steady_state_ms_per_step_l2_warm"],4), "proj ms", round(d["stage_ms"]["projection"],4), "plan", d["plan"]["temporal_block"], d["plan"]["tile_rows_per_warp"], "e2e", round(d["e2e"]["value"]/1e9,2))
PY
done
