#!/usr/bin/env python
"""single-GPU probe of the --fastidious graft search (profiling aid, not a bench): device time of swb200_d1_fastidious on the bench
workload for several chunk counts; every variant must return the same graft candidates.  usage: probe_fast.py [n] [chunks ...]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import helpers  # noqa: E402
from swarm_b200 import Engine, HostDb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chunks = [int(x) for x in sys.argv[2:]] or [4, 8, 16, 32, 64]
fa = f"/dev/shm/swb200_{n}x150_s42.fa"
if not Path(fa).exists():
    helpers.make_fasta(fa, n, 150, 42)
db = HostDb(fa)
eng = Engine(0, collect_stats=1)
eng.load(db)
eng.d1_index(); eng.d1_network(); eng.d1_cluster(want=())
ref = None
for ch in chunks:
    eng.set_option("fast_chunks", ch)
    ms = []
    for _ in range(4):
        gc, nl, nh = eng.d1_fastidious(boundary=3)
        ms.append(round(eng.phase_seconds(4) * 1e3, 3))
    st = eng.stats()
    print("chunks", ch, "fastidious ms", ms[1:], "light", nl, "heavy", nh, {k: st[k] for k in ("fast_light_variants", "fast_heavy_variants", "fast_tag_matches", "fast_verified")}, flush=True)
    if ref is None:
        ref = gc.copy()
    else:
        print("  same graft candidates as the first variant:", bool(np.array_equal(ref, gc)), flush=True)
eng.close()
