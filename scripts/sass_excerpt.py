#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libswarm_b200.so, the count of the mnemonics that matter on this path (UBLKCP = 1-D TMA
bulk copy, SYNCS = mbarrier, LDGSTS = cp.async, ATOMS/ATOMG/RED = atomics, SHFL/VOTE/MATCH = warp collectives) and the instructions
around every UBLKCP.  usage: sass_excerpt.py [kernel substring ...] > profiles/<tag>_sass_excerpts.txt   (no GPU needed)"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "swarm_b200" / "libswarm_b200.so"
want = sys.argv[1:] or ["k_ts_scatter", "k_ts_join", "k_ts_route", "k_ts_scatter_inbox", "k_cluster_persistent", "k_cluster_bucket", "k_fj_", "k_dn_align", "k_derep", "k_d1_network"]
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
MN = ("UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "RED.", "REDG", "SHFL", "VOTE", "MATCH", "POPC", "LDG", "STG", "LDS", "STS", "BAR")
cur, body = None, {}
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        body[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        body[cur].append(line.split("/*", 2)[1][5:].strip() if False else re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
print(f"# cuobjdump -sass {so.name} (sm_100a), CUDA {subprocess.run(['nvcc', '--version'], capture_output=True, text=True).stdout.split('release ')[-1].split(',')[0]}")
for name, ins in body.items():
    if not any(w in name for w in want):
        continue
    c = Counter()
    for i in ins:
        op = i.split("*/")[-1].strip().lstrip("@!UP0123456789 ").split(" ")[0] if "*/" in i else i
        for mn in MN:
            if op.startswith(mn):
                c[mn] += 1
    print(f"\n== {demangle(name)[:150]}\n   {len(ins)} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in c.most_common()))
    for idx, i in enumerate(ins):
        if "UBLKCP" in i:
            for j in range(max(0, idx - 4), min(len(ins), idx + 6)):
                if any(k in ins[j] for k in ("UBLKCP", "SYNCS", "ELECT", "UMOV", "R2UR", "MEMBAR", "FENCE")):
                    print("      " + ins[j][:120])
