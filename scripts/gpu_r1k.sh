#!/bin/bash
# r1k: partitioned tile join + persistent clustering: parity, bench, ncu captures of the two kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_d1.py -m gpu -x -q -k "join or golden_cases or seeded_sets or sharding or large_set or duplicates or edge or mixed or long or cluster_breaking" > $O/r1k_pytest.log 2>&1; tail -5 $O/r1k_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/r1k_bench.json 2> $O/r1k_bench.err; tail -3 $O/r1k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1k_bench.json'))
print(d['value'], d['phases_ms'], d['e2e'], d['counters_per_amplicon'], d['gpu_launches'], d['swarms'])
PY
if [ "$1" = "prof" ]; then
for k in '^k_tile_join$' 'k_cluster_persistent' 'k_tile_partition'; do
  n=$(echo $k | tr -d '^$')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -o $O/prof_r1k_$n -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la $O/*.ncu-rep
fi
