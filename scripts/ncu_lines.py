#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu --set full capture (taken with --import-source on):
the SASS page of the report (executed instructions per SASS instruction) joined with the line table of the cubin
(nvdisasm -g) by instruction index.  usage: ncu_lines.py <ncu-rep> <mangled kernel name substring> [libswarm_b200.so] [top]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, kname = sys.argv[1], sys.argv[2]
so = str(Path(sys.argv[3] if len(sys.argv) > 3 else Path(__file__).resolve().parent.parent / "swarm_b200" / "libswarm_b200.so").resolve())
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ie, si, sm = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
sass = [(r[si].strip(), int(r[ie]), int(r[sm] or 0)) for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith("//---") and kname in l and ".text." in l)
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"# warning: {len(lines)} disassembled instructions vs {len(sass)} in the report (joined by index up to the shorter)")
agg, smp = defaultdict(int), defaultdict(int)
for (f, ln), (_s, n, k) in zip(lines, sass):
    agg[(f, ln)] += n
    smp[(f, ln)] += k
tot, tots = sum(agg.values()), sum(smp.values())
print(f"# {kname}: {tot} warp instructions executed, {tots} stall samples; top {top} source lines")
src = {}
for (f, ln), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        p = Path(__file__).resolve().parent.parent / "swarm_b200" / "csrc" / f
        src[f] = p.read_text().splitlines() if p.exists() else []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print(f"{n:>11} {100 * n / tot:5.1f}%  samples {100 * smp[(f, ln)] / max(tots, 1):5.1f}%  {f}:{ln}  {text}")
