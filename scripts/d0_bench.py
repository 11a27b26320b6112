"""d = 0 (dereplication) on the bench workload shape: device time of swb200_d0_dereplicate (phase 7), algorithmic
bytes and the fraction of the measured HBM peak.  usage: d0_bench.py [unique amplicons] [max copies]
copies = 1: the duplicate-free BASELINE set (every amplicon its own cluster); copies > 1: reads, every sequence present
1..copies times with abundance 1 (what -d 0 is for)."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers
from swarm_b200 import Engine, HostDb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
fa = f"/dev/shm/swb200_{n}x150_s42.fa"
if not Path(fa).exists(): helpers.make_fasta(fa, n, 150, 42)
db = HostDb(fa)
words, lens, ab, stride = db.words, db.len, db.abundance, db.stride
if copies > 1:
    rng = np.random.default_rng(1)
    idx = np.repeat(np.arange(n), rng.integers(1, copies + 1, size=n)); rng.shuffle(idx)
    words = db.words.reshape(n, stride)[idx].copy().reshape(-1); lens = db.len[idx].copy(); ab = np.ones(len(idx), dtype=np.uint64)
eng = Engine(0); eng.set_option("collect_stats", 0); eng.load_db(words, stride, lens, ab)
ms = []
for _ in range(steps + 3):
    rep, mass, size, singles, k = eng.d0_dereplicate(); ms.append(eng.phase_seconds(7) * 1e3)
ms = ms[3:]
rows = len(lens)
t = float(np.mean(ms)) * 1e-3
# algorithmic bytes: packed row + length + abundance read, one 8-byte table entry read and written, slot_of written and
# read, rep written, and per verified match (rows - clusters) the representative's row re-read; the class sums are 16 B
# of atomics per warp-group (counted per row, an upper bound)
bytes_alg = rows * (8 * stride + 4 + 8 + 16 + 8 + 4 + 16) + (rows - k) * 8 * stride
peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6549.1
print(json.dumps({"metric": "sequences dereplicated/s (device-timed) at d=0", "value": rows / t, "unit": "sequences/s", "rows": rows, "clusters": int(k),
                  "ms_per_step": t * 1e3, "ms_min": min(ms), "steps": steps,
                  "roofline": {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_alg / t / 1e9 / peak,
                               "bytes_per_row": bytes_alg / rows, "kernels": "k_derep_claim + k_derep_gather + 4 memsets"}}))
