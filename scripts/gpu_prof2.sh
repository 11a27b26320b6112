#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 900 python bench.py > $O/bench11.json 2> $O/bench11.err; cat $O/bench11.json; tail -2 $O/bench11.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench11_ref.json 2>> $O/bench11.err; cat $O/bench11_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r1h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in k_join_candidates k_join_verify k_bfs_relax k_join_index; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/prof_$k -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la $O/*.ncu-rep
