#!/bin/bash
# scratch: per-step times of the default bench, with and without the NVML clock sampler
for s in 0 1; do
  if [ $s = 1 ]; then export SAYAL_BENCH_NO_SAMPLER=1; fi
  for r in 1 2 3; do
    SAYAL_BENCH_SKIP_STRONG=1 python bench.py --skip-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sampler_off=$s', round(d['ms_per_step'],4), d['step_ms'])"
  done
done
