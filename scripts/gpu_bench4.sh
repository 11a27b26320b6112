#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not dn_ and not fastidious" > $O/pytest_gpu5.log 2>&1; tail -5 $O/pytest_gpu5.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench4_v2.json 2> $O/bench4_v2.err; tail -2 $O/bench4_v2.err
python -c "
import json;d=json.load(open('$O/bench4_v2.json'));print('v2',d['value'],d['phases_ms'],d['e2e']['value'],d['e2e']['ms_per_step'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_d1_network_half -s 1 -c 1 -o $O/prof_network_v2 -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --amplicons 4000000 > $O/ncu_full_v2.out 2>&1
ls -la $O/prof_network_v2.ncu-rep
