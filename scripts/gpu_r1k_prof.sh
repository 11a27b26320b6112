#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
for k in '^k_tile_join$' 'k_cluster_persistent'; do
  n=$(echo $k | tr -d '^$')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $O/prof_r1k_$n -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la $O/*.ncu-rep
