#!/usr/bin/env python
"""bench.py — headline benchmark: amplicons clustered per second at d=1 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--amplicons 10000000] [--length 150]
    python bench.py --impl reference ...        # the reference's own CPU implementation, same metric

A "step" is one pass of the d=1 hot path (index -> network -> cluster) over one synthetic amplicon
set (BASELINE.md §3.2 generator, seed 42).  N=1 workload = BASELINE.json configs[1]: 10 M x 150 bp.
  value : whole-job amplicons/s with the packed database already resident in HBM (device path only)
  e2e   : the same metric through the C ABI with HOST buffers: swb200_load_db (H2D from pinned host
          memory) + index + network + cluster + D2H of the three result arrays, every step
  roofline     : the network kernel, algorithmic bytes per launch (SURVEY.md §8d formula) / CUDA-event time
  cpu_baseline : oracle/_ref/swarm_timed (the unmodified reference + phase timers) on a bounded sample
N>1 (torchrun, one rank per GPU): the seeds of ONE job are sharded across ranks, the directed links are
exchanged with an NCCL all-gather(v), clustering runs replicated; time = max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "amplicons clustered/s (device-timed) at d=1"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the network kernel(s) at 10 M x 150 bp, from the committed
# ncu --set full captures: JOIN = k_join_candidates (2.954+0.320 GB) + k_join_verify (5.897+0.071 GB), profiles/r1h_*;
# HALF (lean kernel) = 2.639+0.028 GB per 4 M seeds scaled to 10 M, profiles/r1d_*
TRAFFIC = {"join": 9.242e9, "half": 6.67e9, "full": None}
UNIT = "amplicons/s"


def env_int(k, d):
    return int(os.environ.get(k, d))


def make_dataset(n, L, seed, path):
    import helpers
    if not Path(path).exists():
        helpers.make_fasta(path, n, L, seed)
    return path


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        self.p = p
        for line in p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            if self.stop:
                break
        p.terminate()

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def finish(self):
        self.stop = True
        time.sleep(0.15)
        try:
            self.p.terminate()
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/swarm_timed = unmodified
    reference sources + a timing-only progress shim) on the host cores, bounded sample of the workload."""
    if rank != 0:
        return
    import helpers  # noqa: F401
    binp = ROOT / "oracle" / "_ref" / "swarm_timed"
    if not binp.exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/swarm_timed was not built (needs /root/reference at build time)"}))
        return
    sample = args.cpu_sample
    fa = make_dataset(sample, args.length, args.seed, f"/dev/shm/swb200_ref_{sample}x{args.length}_s{args.seed}.fa")
    threads = args.cpu_threads or min(os.cpu_count() or 1, 16)
    times = []
    for it in range(args.warmup + args.steps):
        with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as tf:
            pt = tf.name
        env = dict(os.environ, SWARM_PHASE_TIMES=pt)
        subprocess.run([str(binp), "-t", str(threads), "-l", os.devnull, "-o", os.devnull, fa], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ph = {}
        for line in open(pt):
            k, v = line.rstrip("\n").split("\t")
            ph[k.strip().rstrip(":")] = float(v)
        os.unlink(pt)
        t = ph["Hashing sequences"] + ph["Building network"] + ph["Clustering"]
        if it >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.amplicons} x {args.length} bp synthetic amplicons, d=1 (BASELINE configs[1])",
                       "sample": f"first {sample} amplicons of the same seeded generator stream",
                       "phases_timed": "Hashing sequences + Building network + Clustering (reference's own phases)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{sample} x {args.length} bp, -t {threads}, phases hash+network+cluster"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(args):
    binp = ROOT / "oracle" / "_ref" / "swarm_timed"
    if not binp.exists():
        return None
    sample = args.cpu_sample
    fa = make_dataset(sample, args.length, args.seed, f"/dev/shm/swb200_ref_{sample}x{args.length}_s{args.seed}.fa")
    threads = args.cpu_threads or min(os.cpu_count() or 1, 16)
    with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as tf:
        pt = tf.name
    env = dict(os.environ, SWARM_PHASE_TIMES=pt)
    subprocess.run([str(binp), "-t", str(threads), "-l", os.devnull, "-o", os.devnull, fa], check=True, env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ph = {}
    for line in open(pt):
        k, v = line.rstrip("\n").split("\t")
        ph[k.strip().rstrip(":")] = float(v)
    os.unlink(pt)
    t = ph["Hashing sequences"] + ph["Building network"] + ph["Clustering"]
    return {"value": sample / t, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"first {sample} of the workload's amplicons, reference binary -t {threads}, phases hash+network+cluster = {t:.3f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--amplicons", type=int, default=10_000_000)
    ap.add_argument("--length", type=int, default=150)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--enum-mode", type=int, default=2, help="0 full microvariant enumeration, 1 half, 2 pigeonhole join (default)")
    ap.add_argument("--bloom-bytes", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fastidious", action="store_true", help="BASELINE configs[2]: add the --fastidious graft search to every step")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from swarm_b200 import Engine, HostDb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        if not os.environ.get("BENCH_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL_DEBUG=VERSION/INFO prints to stdout; the contract is ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_total = args.amplicons
    fa = f"/dev/shm/swb200_{n_total}x{args.length}_s{args.seed}.fa"
    if rank == 0:
        make_dataset(n_total, args.length, args.seed, fa)
    if world > 1:
        dist.barrier()
    db = HostDb(fa)
    n = db.n
    # pinned host copies: what a host application hands to swb200_load_db
    pw = torch.empty(n * db.stride, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    pl = torch.empty(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    pa = torch.empty(n, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    pw[:] = db.words
    pl[:] = db.len
    pa[:] = db.abundance
    res = {k: torch.empty(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32) for k in ("swarm_of", "generation", "parent")}
    res_gc = torch.empty(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    h2d = pw.nbytes + pl.nbytes + pa.nbytes
    d2h = (4 if args.fastidious else 3) * 4 * n

    eng = Engine(local, enum_mode=args.enum_mode, bloom_bytes_per_slot=args.bloom_bytes, collect_stats=0,
                 shard_rank=rank, shard_world=world)

    from swarm_b200.multi import exchange_engine_links

    def gather_links():
        exchange_engine_links(eng)

    def device_step():
        eng.d1_index()
        eng.d1_network()
        gather_links()
        eng.d1_cluster(want=())
        if args.fastidious:
            eng.d1_fastidious(want=False)

    def e2e_step():
        eng.load_db(pw, db.stride, pl, pa)
        eng.d1_index()
        eng.d1_network()
        gather_links()
        r = eng.d1_cluster(out=res)
        if args.fastidious:
            eng.d1_fastidious(out=res_gc)
        return r

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    eng.load_db(pw, db.stride, pl, pa)
    for _ in range(args.warmup):
        device_step()
    launches0 = eng.stats()["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    phase = {1: [], 2: [], 3: [], 4: []}
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        device_step()
        for p in phase:
            phase[p].append(eng.phase_seconds(p))
    sync_all()
    dt = time.perf_counter() - t0
    launches = eng.stats()["launches"] - launches0
    eng.set_option("collect_stats", 1)      # one extra, untimed pass with the counting kernel variant
    device_step()
    st = eng.stats()
    eng.set_option("collect_stats", 0)
    # e2e: host buffers in, host arrays out, every step
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sw, gen, par = e2e_step()
    sync_all()
    dt_e2e = time.perf_counter() - t0
    clocks = sampler.finish() if rank == 0 else None

    times = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dt, dt_e2e = float(times[0]), float(times[1])

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        net_s = sum(phase[2]) / len(phase[2])
        # algorithmic bytes per amplicon: B1 = P + 16 + 8 V + 12 s + (P+8) c + 4 e with the
        # implementation's own counted V (variants probed), s (slots visited), c (exact compares), e (links)
        seeds = max(1, (n + world - 1) // world)
        P_ = 8 * ((args.length + 31) // 32)
        V, s_, c_, e_ = st["variants"] / seeds, st["slots_visited"] / seeds, st["exact_compares"] / seeds, st["links"] / seeds
        if args.enum_mode == 2:
            # JOIN: own sequence + 2 piece entries written/read (8 B slots visited) + candidate list (8 B written + read)
            # + both packed sequences and abundances per exact comparison + links written
            b1_counted = P_ + 16 + 8 * s_ + 16 * (st["filter_pass"] / seeds) + (2 * P_ + 16) * c_ + 8 * e_
        else:
            b1_counted = P_ + 16 + 8 * V + 16 * s_ + (P_ + 8) * c_ + 8 * e_
        b1_survey = 8400.0 if args.length == 150 else (P_ + 16 + 8 * (7 * args.length + 4))
        achieved = seeds * b1_counted / net_s / 1e9
        line = {
            "metric": METRIC, "value": n * args.steps / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{n} x {args.length} bp synthetic amplicons (seed {args.seed}), d=1" + (" --fastidious, BASELINE configs[2]" if args.fastidious else ", BASELINE configs[1]"),
                       "enum_mode": {0: "full", 1: "half", 2: "join"}[args.enum_mode], "filter_bytes_per_slot": args.bloom_bytes,
                       "l2": "inputs larger than L2 (packed db + table + filter = %.0f MB)" % ((pw.nbytes + 16 * 1.68e7 + 1.68e7) / 1e6),
                       "parallelism": (f"K-mer table sharded by hash range over {world} GPUs; links all-gathered (NCCL all-gather over NVLink); clustering replicated"
                                       if world > 1 else "single GPU")},
            "phases_ms": {"index": 1e3 * sum(phase[1]) / len(phase[1]), "network": 1e3 * net_s,
                          "cluster": 1e3 * sum(phase[3]) / len(phase[3]),
                          "fastidious": (1e3 * sum(phase[4]) / len(phase[4])) if args.fastidious else None},
            "e2e": {"value": n * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * dt_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC.get({0: "full", 1: "half", 2: "join"}[args.enum_mode]) if (n == 10_000_000 and world == 1) else None,
                         "traffic_source": "profiles/r1h_*_full_set_10M.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, network kernels)",
                         "kernel": {0: "k_d1_network<FULL>", 1: "k_d1_network_half", 2: "k_join_candidates + k_join_verify (network phase)"}[args.enum_mode], "bytes_per_amplicon_counted": b1_counted,
                         "bytes_per_amplicon_survey_formula_full_enumeration": b1_survey,
                         "achieved_if_counted_as_full_enumeration": seeds * b1_survey / net_s / 1e9,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "counters_per_amplicon": {"variants": V, "filter_pass": st["filter_pass"] / seeds, "slots_visited": s_,
                                      "exact_compares": c_, "links": e_},
            "swarms": int((sw == np.arange(n, dtype=np.uint32)).sum()),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
