#!/bin/bash
# r1v: d=0 and -u on the GPU, clustering cache-hint experiment, box diagnostics, bench
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
nvidia-smi -q | grep -E -i "product name|ecc mode|current|pending|graphics  |sm  |memory  |mig mode|persistence|power limit|compute mode|bar1|Total  " | head -40 > $O/r1v_smi.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,ecc.mode.current,memory.total,power.limit --format=csv >> $O/r1v_smi.txt
timeout 900 python -m pytest tests/test_derep.py tests/test_cli.py -m gpu -x -q > $O/r1v_pytest.log 2>&1; tail -3 $O/r1v_pytest.log
timeout 900 python scripts/cluster_probe.py > $O/r1v_probe.log 2>&1; cat $O/r1v_probe.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > $O/r1v_bench.json 2> $O/r1v_bench.err; tail -2 $O/r1v_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1v_bench.json')); print('%.4g'%d['value'], d['phases_ms'], 'e2e %.4g'%d['e2e']['value'])
PY
