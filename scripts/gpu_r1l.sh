#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
SWB200_CLUSTER_TS=1 timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 1 > $O/r1l_bench.json 2> $O/r1l_bench.err; grep cluster_csr $O/r1l_bench.err | tail -2
for k in 'k_cluster_csr'; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/prof_r1l_$k -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
