#!/bin/bash
# r1y: full GPU suite after the wide aligner / d=0 limit / CLI context thread; CLI wall time at 10 M (ours only)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r1y_pytest_gpu.log 2>&1; tail -3 $O/r1y_pytest_gpu.log
FA=/dev/shm/swb200_10000000x150_s42.fa
[ -f $FA ] || python -c "
import sys; sys.path.insert(0,'tests'); import helpers; helpers.make_fasta('$FA', 10000000, 150, 42)"
for i in 1 2; do
s=$(date +%s.%N); bin/swarm_b200 -o /dev/shm/mine.o -l /dev/shm/mine.log $FA; rc=$?; e=$(date +%s.%N)
python -c "print('cli 10M wall s', round($e-$s,2), 'rc', $rc)"
done
tail -4 /dev/shm/mine.log
