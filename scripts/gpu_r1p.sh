#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
SWB200_CLUSTER_TS=1 timeout 600 python scripts/dist_world1.py > $O/r1p_world1.log 2>&1; tail -8 $O/r1p_world1.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cluster_dist -s 1 -c 1 -o $O/prof_r1p_k_cluster_dist -f python scripts/dist_world1.py > /dev/null 2>&1
ls -la $O/prof_r1p*
