#!/bin/bash
# r1r: full GPU test suite, default bench (both arms), launch list + full-set captures of the step's kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r1r_pytest_gpu.log 2>&1; tail -4 $O/r1r_pytest_gpu.log
timeout 600 python bench.py > $O/r1r_bench.json 2> $O/r1r_bench.err; tail -2 $O/r1r_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r1r_bench_ref.json 2> $O/r1r_bench_ref.err
timeout 600 python bench.py --fastidious --no-cpu-baseline > $O/r1r_bench_fast.json 2> $O/r1r_bench_fast.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r1r_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in '^k_tile_join$' 'k_cluster_persistent' 'k_tile_partition'; do
  n=$(echo $k | tr -d '^$')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/prof_r1r_$n -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls $O/prof_r1r_*
python - <<'PY'
import json
for f in ('r1r_bench','r1r_bench_ref','r1r_bench_fast'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, '%.4g'%d['value'], d.get('phases_ms'), 'e2e %.4g'%d['e2e']['value'], d.get('clocks'))
PY
