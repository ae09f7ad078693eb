#!/bin/bash
# tools/gpu_round.sh <tag> — one gpurun call: golden fixtures from the reference build, GPU tests, both bench arms,
# ncu launch list + full captures.  Everything lands in gpurun_out/.
set -u
tag=${1:-r1}
mkdir -p gpurun_out
python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/${tag}_golden.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
cat gpurun_out/${tag}_bench_reference.json gpurun_out/${tag}_bench_ours.json
bash profiles/capture.sh ${tag} > gpurun_out/${tag}_capture.log 2>&1
ls gpurun_out
