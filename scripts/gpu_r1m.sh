#!/bin/bash
# r1m: new bench (event timing) at N=1, then the weak-scaling bench at N=2 when two GPUs are visible
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" = "1" ]; then
timeout 900 python bench.py > $O/r1m_bench_n1.json 2> $O/r1m_bench_n1.err; tail -3 $O/r1m_bench_n1.err; cat $O/r1m_bench_n1.json
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 > $O/r1m_bench_n$NG.json 2> $O/r1m_bench_n$NG.err; tail -5 $O/r1m_bench_n$NG.err; cat $O/r1m_bench_n$NG.json
fi
