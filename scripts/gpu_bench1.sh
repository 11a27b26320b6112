#!/bin/bash
# first measurement pass on the B200 box (run under gpurun)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
(nvidia-smi; lscpu | head -20; free -g) > $O/box.txt 2>&1
timeout 900 python bench.py > $O/bench_half_b1.json 2> $O/bench_half_b1.err
tail -c 3000 $O/bench_half_b1.json
for t in 8 16 32 64 128; do
  timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --cpu-threads $t > $O/ref_t$t.json 2>> $O/ref.err
  python -c "import json;d=json.load(open('$O/ref_t$t.json'));print('ref threads',$t,d['value'],d['ms_per_step'])"
done
timeout 600 python bench.py --enum-mode 0 --no-cpu-baseline > $O/bench_full_b1.json 2> $O/bench_full_b1.err
timeout 600 python bench.py --bloom-bytes 2 --no-cpu-baseline > $O/bench_half_b2.json 2> $O/bench_half_b2.err
timeout 600 python bench.py --bloom-bytes 4 --no-cpu-baseline > $O/bench_half_b4.json 2> $O/bench_half_b4.err
timeout 600 python bench.py --enum-mode 0 --bloom-bytes 4 --no-cpu-baseline > $O/bench_full_b4.json 2> $O/bench_full_b4.err
for f in bench_full_b1 bench_half_b2 bench_half_b4 bench_full_b4; do python -c "
import json;d=json.load(open('$O/$f.json'));print('$f',d['value'],d['phases_ms'],d['e2e']['value'],d['counters_per_amplicon'])"; done
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r1.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.out 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_d1_network -s 1 -c 1 -o $O/prof_network_r1 -f \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.out 2>&1
ls -la $O
