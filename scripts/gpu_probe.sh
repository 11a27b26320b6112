#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $O/probe_topo.txt 2>&1
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 scripts/probe_multi.py > $O/probe_multi.log 2>&1
grep -E "all_gather|latency|symm|peer|can_device|via|NVLS|P2P|SHM|Channel 00" $O/probe_multi.log | head -40
