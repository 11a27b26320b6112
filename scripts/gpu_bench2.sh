#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu2.log 2>&1; tail -3 $O/pytest_gpu2.log
timeout 900 python bench.py > $O/bench2_half.json 2> $O/bench2_half.err; cat $O/bench2_half.json
M="dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_write.sum"
for d in 0 1 2 3; do
  SWB200_DEBUG=$d timeout 600 ncu --metrics $M --clock-control none -k regex:k_d1_network -s 1 -c 1 --csv --log-file $O/dbg_$d.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --amplicons 4000000 > $O/dbg_$d.out 2>&1
  echo "== dbg $d"; grep -E "dram__bytes_read.sum|hit_rate|gpu__time|inst_executed|lookup|op_read|op_atom|op_red|op_write" $O/dbg_$d.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
