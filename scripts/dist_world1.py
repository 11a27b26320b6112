"""single-GPU timing probe of the multi-GPU clustering kernel (world = 1: every link is local)"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers
from swarm_b200 import Engine, HostDb
from swarm_b200.ffi import dist_buffer_bytes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
fa = f"/dev/shm/swb200_{n}x150_s42.fa"
if not Path(fa).exists(): helpers.make_fasta(fa, n, 150, 42)
db = HostDb(fa)
eng = Engine(0); eng.load(db); eng.d1_index(); eng.d1_network()
nbytes = dist_buffer_bytes(db.n, 1)
buf = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device="cuda")
eng.dist_setup(0, 1, [buf.data_ptr()], nbytes)
for _ in range(3):
    eng.d1_cluster_dist(None); print("dist ms", eng.phase_seconds(3) * 1e3)
for _ in range(2):
    eng.d1_cluster(want=()); print("single ms", eng.phase_seconds(3) * 1e3)
