#!/bin/bash
# r1q: N GPUs: torchrun parity worker (1M amplicons) + weak-scaling bench, dist and replicated clustering
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import helpers; helpers.make_fasta('/dev/shm/par_1m.fa', 1000000, 150, 91, 0)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tests/dist_worker.py /dev/shm/par_1m.fa 2 > $O/r1q_parity_n$NG.log 2>&1; grep "dist ok" $O/r1q_parity_n$NG.log | wc -l; grep -iE "error|assert" $O/r1q_parity_n$NG.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 > $O/r1q_bench_n$NG.json 2> $O/r1q_bench_n$NG.err; tail -2 $O/r1q_bench_n$NG.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 3 --warmup 2 --multi replicated > $O/r1q_bench_repl_n$NG.json 2> $O/r1q_bench_repl_n$NG.err
python - <<PY
import json
for f in ('$O/r1q_bench_n$NG.json','$O/r1q_bench_repl_n$NG.json'):
    try:
        d=json.load(open(f)); print(d['n_gpus'], '%.3g'%d['value'], '%.2f ms'%d['ms_per_step'], d['phases_ms'], 'e2e %.3g %.2f ms'%(d['e2e']['value'], d['e2e']['ms_per_step']), d['swarms'])
    except Exception as e: print(f, 'failed', e)
PY
