#!/bin/bash
# r1x: compact loader (tests + e2e bench) on one GPU; with more GPUs: dist clustering tests, torchrun parity worker, weak-scaling bench
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" = "1" ]; then
timeout 600 python -m pytest tests/test_gpu_d1.py -m gpu -x -q -k "compact or sharding or golden_cases" > $O/r1x_pytest.log 2>&1; tail -3 $O/r1x_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/r1x_bench.json 2> $O/r1x_bench.err; tail -2 $O/r1x_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1x_bench.json')); print('%.4g'%d['value'], d['phases_ms'], d['e2e'])
PY
else
timeout 400 python -m pytest tests/test_gpu_d1.py -m gpu -x -q -k "dist_clustering or sharding" > $O/r1x_pytest_n$NG.log 2>&1; tail -3 $O/r1x_pytest_n$NG.log
python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import helpers; helpers.make_fasta('/dev/shm/par_1m.fa', 1000000, 150, 91, 0)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tests/dist_worker.py /dev/shm/par_1m.fa 2 > $O/r1x_parity_n$NG.log 2>&1; grep "dist ok" $O/r1x_parity_n$NG.log | wc -l; grep -iE "error|assert" $O/r1x_parity_n$NG.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 > $O/r1x_bench_n$NG.json 2> $O/r1x_bench_n$NG.err; tail -2 $O/r1x_bench_n$NG.err | cut -c1-300
python - <<PY
import json
d=json.load(open('$O/r1x_bench_n$NG.json')); print(d['n_gpus'], '%.3g'%d['value'], '%.2f ms'%d['ms_per_step'], d['phases_ms'], 'e2e %.3g %.2f ms'%(d['e2e']['value'], d['e2e']['ms_per_step']), d['swarms'])
PY
fi
