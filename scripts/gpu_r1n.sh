#!/bin/bash
# r1n: multi-GPU path: dist clustering tests (world 1, world 2 via torchrun), then the weak-scaling bench
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 400 python -m pytest tests/test_gpu_d1.py -m gpu -x -q -k "dist_clustering" > $O/r1n_pytest.log 2>&1; tail -15 $O/r1n_pytest.log
if [ "$NG" != "1" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 2 $BENCH_EXTRA > $O/r1n_bench_n$NG.json 2> $O/r1n_bench_n$NG.err; tail -3 $O/r1n_bench_n$NG.err | cut -c1-300
python - <<PY
import json
d=json.load(open('$O/r1n_bench_n$NG.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['phases_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['swarms'])
PY
fi
