#!/usr/bin/env python
"""single-GPU probe (profiling aid, not a bench): phase times of the d=1 step for several kernel choices on the bench
workload, device-timed per phase.  usage: probe.py [n] [variant ...]   variants: name=opt:val,opt:val"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
from swarm_b200 import Engine, HostDb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
variants = sys.argv[2:] or ["default=", "fat=tile_rows:1", "r1=join_kernel:2", "frontier=cluster_kernel:4"]
fa = f"/dev/shm/swb200_{n}x150_s42.fa"
if not Path(fa).exists():
    helpers.make_fasta(fa, n, 150, 42)
t = time.time(); db = HostDb(fa); print("parse s", round(time.time() - t, 2), flush=True)
ref = None
for v in variants:
    tag, _, spec = v.partition("=")
    opts = {k: int(x) for k, x in (kv.split(":") for kv in spec.split(",") if kv)}
    eng = Engine(0, **opts)
    eng.load(db)
    rows = []
    for it in range(5):
        eng.d1_index(); eng.d1_network(); sw, gen, par = eng.d1_cluster()
        rows.append([round(eng.phase_seconds(p) * 1e3, 3) for p in (1, 2, 3)])
    eng.set_option("collect_stats", 1)
    eng.d1_index(); eng.d1_network(); eng.d1_cluster(want=())
    st = eng.stats()
    print(tag, opts, "index/network/cluster ms:", rows[1:], "links", eng.n_links, "rounds", st["cluster_rounds"],
          {k: round(st[k] / n, 3) for k in ("variants", "filter_pass", "slots_visited", "exact_compares", "rows_gathered")},
          "overflow", st["tile_overflow"], flush=True)
    if ref is None:
        ref = (sw.copy(), gen.copy(), par.copy())
    else:
        import numpy as np
        print("  same result as the first variant:", all(np.array_equal(a, b) for a, b in zip(ref, (sw, gen, par))), flush=True)
    eng.close()
