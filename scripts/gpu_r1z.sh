#!/bin/bash
# r1z: full GPU suite (wide-d related-sequence tests) + default bench with per-phase roofline and two-jobs-in-flight e2e
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r1z_pytest_gpu.log 2>&1; tail -3 $O/r1z_pytest_gpu.log
timeout 600 python bench.py > $O/r1z_bench.json 2> $O/r1z_bench.err; tail -2 $O/r1z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1z_bench.json')); print('%.4g'%d['value'], d['phases_ms'], d['e2e'], d['roofline'].get('phases'), d['cpu_baseline'])
PY
