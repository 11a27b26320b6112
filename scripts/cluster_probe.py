"""single-GPU probe: clustering kernel variants on the bench workload + d=0 timing (profiling aid, not a bench)"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers
from swarm_b200 import Engine, HostDb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
fa = f"/dev/shm/swb200_{n}x150_s42.fa"
if not Path(fa).exists(): helpers.make_fasta(fa, n, 150, 42)
t = time.time(); db = HostDb(fa); print("parse s", round(time.time() - t, 2), flush=True)
eng = Engine(0); eng.load(db); eng.d1_index(); eng.d1_network()
for tag, opts in (("hints0", {"cluster_kernel": 0, "cluster_hints": 0}), ("hints1", {"cluster_kernel": 0, "cluster_hints": 1}),
                  ("csr", {"cluster_kernel": 3}), ("hints0", {"cluster_kernel": 0, "cluster_hints": 0}), ("hints1", {"cluster_kernel": 0, "cluster_hints": 1})):
    for k, v in opts.items(): eng.set_option(k, v)
    ms = []
    for _ in range(4):
        eng.d1_cluster(want=()); ms.append(round(eng.phase_seconds(3) * 1e3, 3))
    print("cluster", tag, ms, flush=True)
# d = 0 on the same (duplicate-free) database, then on reads with duplicates: every sequence 1..8 times
for _ in range(3):
    rep, mass, size, singles, k = eng.d0_dereplicate(); print("d0 unique ms", round(eng.phase_seconds(7) * 1e3, 3), k, flush=True)
m = 2_000_000
rng = np.random.default_rng(1)
idx = np.repeat(np.arange(m), rng.integers(1, 9, size=m)); rng.shuffle(idx)
words = db.words.reshape(db.n, db.stride)[idx].copy().reshape(-1)
rlen = db.len[idx].copy()
eng2 = Engine(0); eng2.load_db(words, db.stride, rlen, np.ones(len(idx), dtype=np.uint64))
class rd: n = len(idx)
for _ in range(3):
    rep, mass, size, singles, k = eng2.d0_dereplicate(); print("d0 reads ms", round(eng2.phase_seconds(7) * 1e3, 3), rd.n, k, flush=True)
assert k == m and int(size.sum()) == rd.n and np.array_equal(rep[rep], rep)
first = np.full(m, rd.n, dtype=np.int64); np.minimum.at(first, idx, np.arange(rd.n)); assert np.array_equal(rep, first[idx])
print("d0 reads parity ok")
