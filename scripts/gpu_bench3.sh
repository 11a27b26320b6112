#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu3.log 2>&1; tail -15 $O/pytest_gpu3.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench3_half.json 2> $O/bench3_half.err; cat $O/bench3_half.json; tail -3 $O/bench3_half.err
timeout 900 python bench.py --fastidious --steps 2 --warmup 1 --no-cpu-baseline > $O/bench3_fast.json 2> $O/bench3_fast.err; cat $O/bench3_fast.json; tail -3 $O/bench3_fast.err
