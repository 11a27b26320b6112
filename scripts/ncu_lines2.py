#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel straight from an ncu --set full report taken with
--import-source on: the combined CUDA + SASS source page already carries the line of every SASS instruction.
usage: ncu_lines2.py <ncu-rep> [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg, smp, text = defaultdict(int), defaultdict(int), {}
fname, hdr = "?", None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():      # rows without a line number are the SASS instructions themselves
        continue
    try:
        ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
        n = int(r[ie]) if r[ie].isdigit() else 0
        k = int(r[sm]) if r[sm].isdigit() else 0
    except (ValueError, IndexError):
        continue
    key = (fname, r[0])
    agg[key] += n
    smp[key] += k
    text.setdefault(key, r[1].strip()[:110])
tot, tots = sum(agg.values()), sum(smp.values())
print(f"# {rep}: {tot} warp instructions executed, {tots} stall samples; top {top} source lines (file:line, share of instructions, share of samples)")
for key, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{n:>11} {100 * n / max(tot, 1):5.1f}%  samples {100 * smp[key] / max(tots, 1):5.1f}%  {key[0]}:{key[1]}  {text[key]}")
