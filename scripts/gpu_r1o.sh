#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
for dbg in 0 1; do
SWB200_DIST_DBG=$dbg SWB200_CLUSTER_TS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 1 > $O/r1o_bench_dbg$dbg.json 2> $O/r1o_bench_dbg$dbg.err
echo "dbg=$dbg"; grep "cluster_dist rank 1" $O/r1o_bench_dbg$dbg.err | tail -1 | cut -c1-700
done
