#!/bin/bash
# r1j: full GPU test suite + default bench (both arms) on a fresh box
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/r1j_smi.txt; nproc >> $O/r1j_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r1j_pytest_gpu.log 2>&1; tail -5 $O/r1j_pytest_gpu.log
timeout 600 python bench.py > $O/r1j_bench.json 2> $O/r1j_bench.err; tail -3 $O/r1j_bench.err; cat $O/r1j_bench.json
timeout 600 python bench.py --impl reference > $O/r1j_bench_ref.json 2> $O/r1j_bench_ref.err; cat $O/r1j_bench_ref.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
