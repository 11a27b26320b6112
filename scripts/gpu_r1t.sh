#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_tile_join$' -s 1 -c 1 -o $O/prof_r1t_k_tile_join -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la $O/prof_r1t*
