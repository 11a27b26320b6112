#!/usr/bin/env python
"""BASELINE configs[3]: N x 400 bp synthetic amplicons, d=2 — device time of swb200_dn_cluster, and the
reference binary on a bounded sample.  usage: bench_dn.py [N] [ref_sample]"""
import json, os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import helpers
from swarm_b200 import Engine, HostDb, scoring

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref_n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
fa = f"/dev/shm/swb200_dn_{n}x400.fa"
if not os.path.exists(fa):
    helpers.make_fasta(fa, n, 400, 42)
t0 = time.time(); db = HostDb(fa, check_dup_sequences=True); t_parse = time.time() - t0
out = {"workload": f"{n} x 400 bp, d=2", "parse_s": t_parse}
for filt in (0, 1) if n <= 200_000 else (0,):
    eng = Engine(0, dn_filter=filt)
    eng.load(db)
    ts = []
    for _ in range(3):
        sw, gen, par, pd = eng.dn_cluster(2, penalties=scoring())
        ts.append(eng.phase_seconds(6))
    st = eng.stats()
    out[f"filter{filt}"] = {"device_s": min(ts), "amplicons_per_s": n / min(ts), "tasks": st["dn_alignments"], "pruned": st["dn_pruned"],
                             "links": st["dn_links"], "qgram_cmp": st["dn_qgram_comparisons"], "swarms": int((sw == __import__('numpy').arange(n)).sum())}
    eng.close()
if helpers.have_ref() and ref_n:
    rfa = f"/dev/shm/swb200_dn_{ref_n}x400.fa"
    if not os.path.exists(rfa):
        helpers.make_fasta(rfa, ref_n, 400, 42)
    with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as tf:
        pt = tf.name
    t0 = time.time()
    subprocess.run([str(ROOT / "oracle/_ref/swarm_timed"), "-d", "2", "-t", "16", "-l", os.devnull, "-o", os.devnull, rfa], check=True,
                   env=dict(os.environ, SWARM_PHASE_TIMES=pt), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ph = dict(l.rstrip("\n").split("\t") for l in open(pt))
    t = sum(float(v) for k, v in ph.items() if k.strip().startswith(("Find qgram", "Clustering")))
    out["reference"] = {"sample": ref_n, "threads": 16, "phases_s": t, "amplicons_per_s": ref_n / t, "wall_s": time.time() - t0}
print(json.dumps(out))
