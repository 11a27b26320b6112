#!/bin/bash
# r1w: full GPU suite, smoke, both bench arms, launch list, d=0 bench + full-set captures of its kernels, CLI wall time
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r1w_pytest_gpu.log 2>&1; tail -3 $O/r1w_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > $O/r1w_bench.json 2> $O/r1w_bench.err; tail -2 $O/r1w_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r1w_bench_ref.json 2> $O/r1w_bench_ref.err
timeout 300 python scripts/d0_bench.py 10000000 1 > $O/r1w_d0_unique.json 2> $O/r1w_d0.err; cat $O/r1w_d0_unique.json | cut -c1-400
timeout 300 python scripts/d0_bench.py 2500000 7 > $O/r1w_d0_reads.json 2>> $O/r1w_d0.err; cat $O/r1w_d0_reads.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in k_derep_claim k_derep_gather; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 3 -c 1 -o $O/prof_r1w_$k -f python scripts/d0_bench.py 10000000 1 2 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cluster_persistent -s 1 -c 1 -o $O/prof_r1w_k_cluster_persistent -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls $O/prof_r1w_* 2>&1 | head
timeout 900 bash scripts/cli_wall.sh 10000000 > $O/r1w_cli_wall.json 2> $O/r1w_cli_wall.err; cat $O/r1w_cli_wall.json
python - <<'PY'
import json
for f in ('r1w_bench','r1w_bench_ref'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, '%.4g'%d['value'], d.get('phases_ms'), 'e2e %.4g'%d['e2e']['value'], d.get('clocks'))
PY
