#!/bin/bash
# One parametrised runner for the GPU box (under gpurun): scripts/gpu_run.sh TAG step [step ...]
#   steps: box | full:<kernel regex>[:skip[:bench args[:mangled-name substring for the source-line table]]] | fullp:<kernel regex>[:skip] (ncu --set full of one launch under scripts/probe.py) | tests[:pytest -k expr] | scale | bench[:extra args] | probe[:args] | launches | full:<kernel regex>[:skip] | sanitize | ts
# Everything lands in gpurun_out/TAG_*.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
TAG=$1; shift
O=gpurun_out
for step in "$@"; do
  name=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  case $name in
    box) (nvidia-smi; nvidia-smi topo -m; lscpu | head -20; free -g) > $O/${TAG}_box.txt 2>&1 ;;
    tests) if [ -n "$arg" ]; then timeout 1500 python -m pytest tests -x -q -m gpu -k "$arg" > $O/${TAG}_tests.log 2>&1; else timeout 2400 python -m pytest tests -x -q -m gpu > $O/${TAG}_tests.log 2>&1; fi
           echo "== tests rc=$?"; tail -5 $O/${TAG}_tests.log ;;
    scale) timeout 1500 python -m pytest tests/test_gpu_scale.py -x -q -m gpu > $O/${TAG}_scale.log 2>&1; echo "== scale rc=$?"; tail -5 $O/${TAG}_scale.log ;;
    bench) n=$(echo "$arg" | tr -c 'a-zA-Z0-9' '_'); timeout 1500 python bench.py $arg > $O/${TAG}_bench$n.json 2> $O/${TAG}_bench$n.err; echo "== bench $arg rc=$?"
           python -c "import json,sys; d=json.loads(open('$O/${TAG}_bench$n.json').read().strip().splitlines()[-1]); print({k: d.get(k) for k in ('impl','value','ms_per_step','phases_ms','parity','unavailable')}, 'e2e', d.get('e2e',{}).get('ms_per_step'), 'roofline', {k: (d.get('roofline') or {}).get(k) for k in ('kernel','frac')}, 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; tail -3 $O/${TAG}_bench$n.err ;;
    probe) timeout 900 python scripts/probe.py $arg > $O/${TAG}_probe.log 2>&1; echo "== probe rc=$?"; cat $O/${TAG}_probe.log ;;
    ts) SWB200_CLUSTER_TS=1 timeout 600 python scripts/probe.py 10000000 default= > $O/${TAG}_ts.log 2>&1; grep -E "cluster_frontier|cluster_dist" $O/${TAG}_ts.log | tail -3 ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv \
                python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/${TAG}_launches.out 2>&1
              python scripts/ncu_summary.py launches $O/${TAG}_launches.csv > $O/${TAG}_launches.txt; head -20 $O/${TAG}_launches.txt ;;
    full) IFS=: read -r k skip extra mangled <<< "$arg"; skip=${skip:-0}; mangled=${mangled:-$k}
          timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/${TAG}_$k -f \
            python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity $extra > $O/${TAG}_full_$k.out 2>&1
          python scripts/ncu_summary.py full $O/${TAG}_$k.ncu-rep > $O/${TAG}_$k.txt 2>&1
          python scripts/ncu_lines2.py $O/${TAG}_$k.ncu-rep 40 >> $O/${TAG}_$k.txt 2>&1
          [ -n "$KEEP_REP" ] || rm -f $O/${TAG}_$k.ncu-rep      # gpurun brings back at most 64 MiB: the summaries travel, the reports do not
          head -50 $O/${TAG}_$k.txt ;;
    fullp) k=${arg%%:*}; skip=0; [[ "$arg" == *:* ]] && skip=${arg#*:}
          timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/${TAG}_$k -f \
            python scripts/probe.py 10000000 default= > $O/${TAG}_fullp_$k.out 2>&1
          python scripts/ncu_summary.py full $O/${TAG}_$k.ncu-rep > $O/${TAG}_$k.txt 2>&1; head -45 $O/${TAG}_$k.txt ;;
    san) timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests -x -q -m gpu -k "$arg" > $O/${TAG}_san.log 2>&1
         grep -E "=========|passed|failed" $O/${TAG}_san.log | head -60 ;;
    mbench) g=${arg%%:*}; extra=""; [[ "$arg" == *:* ]] && extra=${arg#*:}
            SWB200_CLUSTER_TS=${TS:-} timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $g $extra > $O/${TAG}_mbench_$g.json 2> $O/${TAG}_mbench_$g.err
            echo "== mbench $g rc=$?"; tail -c 3000 $O/${TAG}_mbench_$g.json; grep -E "cluster_dist rank 0|Error|error" $O/${TAG}_mbench_$g.err | tail -5 ;;
    sanitize) bash scripts/gpu_sanitize.sh ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la $O | tail -20
