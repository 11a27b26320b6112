#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r1u_pytest_gpu.log 2>&1; tail -3 $O/r1u_pytest_gpu.log
SWB200_CLUSTER_TS=1 timeout 600 python scripts/dist_world1.py 2>&1 | tail -4 | cut -c1-500
timeout 600 python bench.py --no-cpu-baseline > $O/r1u_bench.json 2> $O/r1u_bench.err; tail -2 $O/r1u_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1u_bench.json')); print('%.4g'%d['value'], d['phases_ms'], 'e2e %.4g'%d['e2e']['value'])
PY
