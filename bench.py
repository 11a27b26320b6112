#!/usr/bin/env python
"""bench.py — headline benchmark: amplicons clustered per second (BASELINE.json), CUDA engine vs the reference's CPU code.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--amplicons N] [--length L]
    python bench.py --impl reference ...        # the reference's own CPU implementation, same metric, same config

A "step" is one pass of the hot path over one synthetic amplicon set (BASELINE.md §3.2 generator, seed 42):
  c2 (default)  10 M x 150 bp, d=1: index -> network -> cluster                      BASELINE configs[1]
  c3            the same + the --fastidious graft search                              configs[2]
  c4            1 M x 400 bp, d=2: q-gram vectors, candidates, exact aligner, cluster configs[3]
  c5            --gpus 8: 8 x 12.5 M = 100 M x 150 bp, d=1, ONE job                   configs[4]
  value : whole-job amplicons/s with the packed database already resident in HBM (device path only)
  e2e   : the same metric through the C ABI with HOST buffers: upload from pinned host memory + the step + D2H of the result
          arrays into pinned host memory, every step
  roofline     : the dominant kernel of the step (largest CUDA-event time): algorithmic bytes per launch / its time; all
                 phases under roofline.phases
  cpu_baseline : oracle/_ref/swarm_timed (the unmodified reference + phase timers) on the host cores
  parity       : UNTIMED leg.  N=1: the engine's -o/-s/-i texts, hashed, against the committed digests of the reference's
                 outputs on the same input (tests/golden/scale_hashes.json); N>1: every rank's rows against the single-GPU
                 engine run on the gathered database by rank 0.
N>1 (torchrun, one rank per GPU), weak scaling: ONE clustering job of N x --amplicons amplicons (see swarm_b200/multi.py).
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

UNIT = "amplicons/s"
CONFIGS = {
    # name: (amplicons per GPU, length, d, fastidious, BASELINE label, scale-hash case, metric)
    "c2": (10_000_000, 150, 1, False, "BASELINE configs[1]", "c2", "amplicons clustered/s (device-timed) at d=1"),
    "c3": (10_000_000, 150, 1, True, "BASELINE configs[2]", "c3", "amplicons clustered/s (device-timed) at d=1 --fastidious"),
    "c4": (1_000_000, 400, 2, False, "BASELINE configs[3]", "c4", "amplicons clustered/s (device-timed) at d=2"),
    "c5": (12_500_000, 150, 1, False, "BASELINE configs[4]", None, "amplicons clustered/s (device-timed) at d=1"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 10 M x 150 bp from the committed ncu --set full captures
TRAFFIC = {}          # kernel -> (bytes, profiles/ file); filled in from profiles/traffic.json when present
JSON_FD = 1


def emit(line):
    """the one JSON line of the contract, on the real stdout"""
    sys.stdout.flush()
    os.write(JSON_FD, (json.dumps(line) + "\n").encode())


def env_int(k, d):
    return int(os.environ.get(k, d))


def make_dataset(n, L, seed, path):
    import helpers
    if not Path(path).exists():
        helpers.make_fasta(path + ".tmp%d" % os.getpid(), n, L, seed)
        os.replace(path + ".tmp%d" % os.getpid(), path)
    return path


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = None
        self.p = None

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


# ---- the reference's CPU implementation (oracle/_ref/swarm_timed = unmodified reference sources + a timing-only progress shim)
REF_PHASES = {
    1: ("Hashing sequences", "Building network", "Clustering"),
    "f": ("Hashing sequences", "Building network", "Clustering", "Adding light swarm amplicons to Bloom filter",
          "Checking heavy swarm amplicons against Bloom filter", "Grafting light swarms on heavy swarms"),
    2: ("Find qgram vects", "Clustering"),
}


def run_reference(fa, d, fastidious, threads):
    """one run of the reference binary; returns (sum of the hot-path phases in s, per-phase dict)"""
    binp = ROOT / "oracle" / "_ref" / "swarm_timed"
    with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as tf:
        pt = tf.name
    cmd = [str(binp), "-t", str(threads), "-l", os.devnull, "-o", os.devnull]
    if d != 1:
        cmd += ["-d", str(d)]
    if fastidious:
        cmd += ["-f"]
    subprocess.run(cmd + [fa], check=True, env=dict(os.environ, SWARM_PHASE_TIMES=pt), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ph = {}
    for line in open(pt):
        k, v = line.rstrip("\n").split("\t")
        ph[k.strip().rstrip(":")] = ph.get(k.strip().rstrip(":"), 0.0) + float(v)
    os.unlink(pt)
    want = REF_PHASES["f" if fastidious else (1 if d == 1 else 2)]
    used = {k: v for k, v in ph.items() if any(k.startswith(w) for w in want)}
    return sum(used.values()), used


def reference_sample(args, cfg, full):
    """(sample size, same_config): the whole workload when one run takes a few tens of seconds, else a bounded prefix of
    the same generator stream"""
    n_gpu, L, d, fast, *_ = cfg
    n = args.amplicons
    if args.cpu_sample:
        return min(args.cpu_sample, n), args.cpu_sample >= n
    if d == 1 and not fast:
        cap = 10_000_000                                  # ~5 s of phases, ~20 s whole process at 16 threads
    elif d == 1:
        cap = 10_000_000 if full else 1_000_000           # -f: a minute or more at 10 M
    else:
        cap = 100_000                                     # d=2 at 1 M x 400 bp: ~20 min
    return min(cap, n), cap >= n


def reference_arm(args, cfg, rank, world):
    """--impl reference: the reference's own CPU code on the host cores, same metric and config.  Rank 0 only."""
    if rank != 0:
        return
    n_gpu, L, d, fast, label, _case, metric = cfg
    binp = ROOT / "oracle" / "_ref" / "swarm_timed"
    if not binp.exists():
        emit({"impl": "reference", "unavailable": "oracle/_ref/swarm_timed was not built (needs /root/reference at build time)"})
        return
    sample, same = reference_sample(args, cfg, True)
    if world > 1:                                         # our arm clusters world x --amplicons as ONE job: the reference runs a bounded sample of it
        same = False
    fa = make_dataset(sample, L, args.seed, f"/dev/shm/swb200_{sample}x{L}_s{args.seed}.fa")
    threads = args.cpu_threads or min(os.cpu_count() or 1, 16)
    # bounded repeats: one run of the full 10 M workload is ~20 s of wall time (FASTA parse + sort included, untimed)
    warm, runs = min(args.warmup, 1), max(1, min(args.steps, 3))
    if fast or d != 1:
        warm, runs = 0, 1
    times, phases = [], None
    for it in range(warm + runs):
        t, phases = run_reference(fa, d, fast, threads)
        if it >= warm:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms * 1e-3)
    wl = f"{args.amplicons * world} x {L} bp synthetic amplicons, d={d}" + (" --fastidious" if fast else "") + f" ({label})"
    emit({"impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": world, "steps": runs, "warmup": warm,
          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
          "config": {"workload": wl, "same_config": bool(same),
                     "sample": ("the whole workload" if same else f"first {sample} amplicons of the same seeded generator stream"),
                     "phases_timed": " + ".join(phases) + " (the reference's own phases, steady_clock)",
                     "requested_steps_warmup": [args.steps, args.warmup]},
          "phases_s": phases,
          "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                           "sample": f"{sample} x {L} bp, -t {threads}, {runs} run(s)"},
          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def cpu_baseline(args, cfg):
    n_gpu, L, d, fast, *_ = cfg
    if not (ROOT / "oracle" / "_ref" / "swarm_timed").exists():
        return None
    sample, same = reference_sample(args, cfg, False)
    fa = make_dataset(sample, L, args.seed, f"/dev/shm/swb200_{sample}x{L}_s{args.seed}.fa")
    threads = args.cpu_threads or min(os.cpu_count() or 1, 16)
    t, phases = run_reference(fa, d, fast, threads)
    return {"value": sample / t, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": ("the whole workload" if same else f"first {sample} of the workload's amplicons") +
                      f", reference binary -t {threads}, phases {' + '.join(phases)} = {t:.3f} s"}


def load_traffic():
    p = ROOT / "profiles" / "traffic.json"
    return json.loads(p.read_text()) if p.exists() else {}


def roofline_block(peak, peaks_found, phases, traffic_key):
    """phases: {name: {"kernel": .., "bytes": algorithmic bytes per launch, "s": CUDA-event seconds, ...}} -> roofline object for
    the phase with the largest time, all phases attached"""
    out = {}
    for k, v in phases.items():
        if v["s"] > 0:
            out[k] = dict(v, achieved=v["bytes"] / v["s"] / 1e9, frac=v["bytes"] / v["s"] / 1e9 / peak, ms=1e3 * v["s"])
            out[k].pop("s")
    top = max(out, key=lambda k: out[k]["ms"])
    tr = load_traffic().get(traffic_key + ":" + top) if traffic_key else None
    return {"bound": "hbm", "achieved": out[top]["achieved"], "peak": peak, "unit": "GB/s", "frac": out[top]["frac"],
            "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
            "kernel": out[top]["kernel"], "phase": top, "bytes_per_launch": out[top]["bytes"],
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks_found else "fallback 6650 GB/s", "phases": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="BASELINE configuration (default c2; c5 needs --gpus 8)")
    ap.add_argument("--amplicons", type=int, default=0, help="amplicons per GPU (the job has gpus x this many); default: the config's")
    ap.add_argument("--length", type=int, default=0)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--enum-mode", type=int, default=2, help="0 full microvariant enumeration, 1 half, 2 pigeonhole join (default)")
    ap.add_argument("--join-kernel", type=int, default=0, help="JOIN: 0 tile store (default), 1 global hash multimap, 2 r1 tile join")
    ap.add_argument("--cluster-kernel", type=int, default=0)
    ap.add_argument("--multi", default="auto", choices=["auto", "replicated", "sharded"],
                    help="N>1 database layout: replicated on every GPU (default for c2) or sharded by rows (default for c5); see bench_multi.py")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity leg (profiling runs)")
    ap.add_argument("--fastidious", action="store_true", help="alias of --config c3")
    args = ap.parse_args()
    if args.config is None:
        args.config = "c3" if args.fastidious else "c2"
    cfg = CONFIGS[args.config]
    args.amplicons = args.amplicons or cfg[0]
    args.length = args.length or cfg[1]
    cfg = (args.amplicons, args.length) + cfg[2:]
    n_gpu, L, d, fast, label, case, metric = cfg

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # the contract is ONE JSON line on stdout: everything any library prints there meanwhile (NCCL's banner / NCCL_DEBUG lines ...)
    # goes to stderr, where the driver can still read it
    global JSON_FD
    sys.stdout.flush()
    JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args, cfg, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import helpers
    from swarm_b200 import D1Result, DnResult, Engine, HostDb, scoring
    from swarm_b200.ffi import compact_form

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if fast or d != 1:
            raise SystemExit("bench.py: c3 and c4 are single-GPU configurations (BASELINE configs[2], [3])")
    if args.config == "c5" and world != 8:
        raise SystemExit("bench.py: c5 is 8 x 12.5 M amplicons: launch with --gpus 8 under torchrun")

    def pinned(count, dtype):
        return torch.empty(count, dtype=dtype, pin_memory=True).numpy()

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    P_ = 8 * ((L + 31) // 32)

    if world > 1:
        from bench_multi import run_multi
        run_multi(args, cfg, rank, world, local, emit, ClockSampler, peak, bool(peaks))
        return

    # ---------------------------------------------------------------- single GPU
    fa = f"/dev/shm/swb200_{args.amplicons}x{L}_s{args.seed}.fa"
    make_dataset(args.amplicons, L, args.seed, fa)
    db = HostDb(fa, check_dup_sequences=d > 1)
    n, stride = db.n, db.stride
    eng = Engine(local, enum_mode=args.enum_mode, join_kernel=args.join_kernel, cluster_kernel=args.cluster_kernel, collect_stats=0)
    # pinned host copies: what a host application hands to the C ABI
    pw, pl, pa = pinned(n * stride, torch.int64).view(np.uint64), pinned(n, torch.int32).view(np.uint32), pinned(n, torch.int64).view(np.uint64)
    pw[:], pl[:], pa[:] = db.words, db.len, db.abundance
    eng.load_db(pw, stride, pl, pa)
    # the compact form of the same database (u16 lengths, abundance runs): what the end-to-end path uploads
    l16, rab, rst = compact_form(pl, pa)
    pl16, prab, prst = pinned(n, torch.int16).view(np.uint16), pinned(rab.shape[0], torch.int64).view(np.uint64), pinned(rst.shape[0], torch.int32).view(np.uint32)
    pl16[:], prab[:], prst[:] = l16, rab, rst
    keys = ("swarm_of", "generation", "parent")
    res = {k: pinned(n, torch.int32).view(np.uint32) for k in keys}
    res_x = pinned(n, torch.int32).view(np.uint32)             # graft candidates (c3) / differences to the parent (c4)
    h2d = pw.nbytes + pl16.nbytes + prab.nbytes + prst.nbytes
    d2h = (4 if (fast or d != 1) else 3) * 4 * n
    pen = scoring()
    from swarm_b200.multi import engine_stream
    ext = engine_stream(eng)

    def device_step():
        if d == 1:
            eng.d1_index()
            eng.d1_network()
            eng.d1_cluster(want=())
            if fast:
                eng.d1_fastidious(want=False)
        else:
            eng.dn_cluster(d, penalties=pen, want=False)

    def e2e_step():
        eng.load_db_compact(pw, stride, pl16, prab, prst)
        if d == 1:
            eng.d1_index()
            eng.d1_network()
            eng.d1_cluster(out=res)
            if fast:
                eng.d1_fastidious(out=res_x)
        else:
            eng.dn_cluster(d, penalties=pen, out=dict(res, pdiff=res_x))

    def timed(step, k):
        """K steps bracketed by synchronize on both sides; device time from CUDA events on the engine's stream"""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record(ext)
        for _ in range(k):
            step()
        ev1.record(ext)
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - t0

    for _ in range(args.warmup):
        device_step()
    launches0 = eng.stats()["launches"]
    sampler = ClockSampler(local)
    sampler.start()
    phase_ids = (1, 2, 3, 4) if d == 1 else (6,)
    phase = {p: [] for p in phase_ids}

    def device_step_logged():
        device_step()
        for p in phase:
            phase[p].append(eng.phase_seconds(p))

    dt, wall = timed(device_step_logged, args.steps)
    launches = eng.stats()["launches"] - launches0
    eng.set_option("collect_stats", 1)      # one extra, untimed pass with the counting kernel variants
    device_step()
    st = eng.stats()
    eng.set_option("collect_stats", 0)
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    dt_e2e, wall_e2e = timed(e2e_step, args.steps)
    clocks = sampler.finish()

    # ---- parity (untimed): the engine's output files, hashed, vs the reference's digests for this input
    parity = None
    if not args.no_parity:
        hashes = helpers.scale_hashes()
        cand = [k for k, v in hashes.items() if (v["n"], v["length"], v["seed"], v["ab_mode"]) == (n, L, args.seed, 0)
                and v["flags"] == " ".join((["-f"] if fast else []) + (["-d", str(d)] if d != 1 else []))]
        if d == 1:
            r = D1Result(db, res["swarm_of"], res["generation"], res["parent"], graft_cand=res_x if fast else None, boundary=3)
        else:
            r = DnResult(db, res["swarm_of"], res["generation"], res["parent"], res_x)
        got = helpers.output_hashes(r.swarms_text(), r.stats_text(), r.structure_text())
        r.close()
        if cand:
            diff = helpers.compare_case(cand[0], got)
            parity = {"ok": diff == [], "vs": f"tests/golden/scale_hashes.json[{cand[0]}]: SHA-256 of the reference's -o (raw + canonical, BASELINE.md §3.4), -s and -i files "
                                              f"on the same FASTA ({hashes[cand[0]]['reference']}, -t {hashes[cand[0]]['reference_threads']})",
                      "differs": diff, "swarms": got["swarms"], "o_canonical_sha256": got["o_canonical_sha256"],
                      "from": "the host arrays of the last end-to-end step"}
        else:
            parity = {"ok": None, "vs": "no committed reference digest for this input size", "swarms": got["swarms"],
                      "o_canonical_sha256": got["o_canonical_sha256"]}

    # ---- roofline: algorithmic bytes per launch of every phase, time = CUDA events inside the timed region
    cnt = {k: st[k] for k in st}
    m_links = eng.n_links if d == 1 else st["dn_links"]
    avg = {p: sum(v) / len(v) for p, v in phase.items()}
    if d == 1:
        e_ = m_links
        rows_ = st["rows_gathered"]
        ph = {
            "index": {"kernel": "k_ts_scatter", "s": avg[1], "bytes": n * (P_ + 12) + 2 * n * 8,
                      "formula": "per amplicon: packed row + length + abundance read (one TMA-staged pass), two 8-byte tile entries written"},
            "network": {"kernel": "k_ts_join (+ k_ts_big, idle on this data)", "s": avg[2], "bytes": 2 * n * 8 + rows_ * P_ + 8 * e_,
                        "formula": "every tile entry read once by TMA, one packed row per entry that has a bucket mate (counted), links written",
                        "rows_gathered_per_amplicon": rows_ / n},
            "cluster": {"kernel": "k_cluster_persistent", "s": avg[3], "bytes": n * 28 + st["cluster_rounds"] * 8 * e_, "rounds": st["cluster_rounds"],
                        "formula": "packed relaxation word initialised (8 B), label + generation + parent written (12 B) and the word read once more for that (8 B) "
                                   "per amplicon; the 8-byte link list re-read every round; relaxation traffic (two random words per ACTIVE link) "
                                   "not counted: a lower bound"},
        }
        if args.join_kernel != 0 or args.enum_mode != 2 or args.cluster_kernel != 0:
            for k in ph:
                ph[k]["kernel"] += " (non-default kernel options: byte formula of the default kernels)"
        if fast:
            ph["fastidious"] = {"kernel": "k_fj_insert + k_fj_candidates + k_fj_verify", "s": avg[4],
                                "bytes": 8 * st["fast_light_variants"] + 32 * st["fast_heavy_variants"] + (2 * P_ + 8) * st["fast_tag_matches"],
                                "formula": "8 B per stored light piece + one 32-byte bucket per heavy lookup + two rows per candidate (counted)"}
        step_bytes = sum(v["bytes"] for v in ph.values())
    else:
        ph = {"dn": {"kernel": "k_dn_qgrams + k_dn_candidates_join + k_dn_align + clustering", "s": avg[6],
                     "bytes": 128 * st["dn_qgram_comparisons"] + 2 * P_ * st["dn_alignments"],
                     "formula": "128 B x q-gram comparisons + (P_q + P_t) x alignments (SURVEY.md §8d), both counted"}}
        step_bytes = ph["dn"]["bytes"]
    roof = roofline_block(peak, bool(peaks), ph, args.config if n == CONFIGS[args.config][0] else None)
    roof["step"] = {"bytes": step_bytes, "achieved": step_bytes / (dt / args.steps) / 1e9, "frac": step_bytes / (dt / args.steps) / 1e9 / peak}

    wl = f"{n} x {L} bp synthetic amplicons, d={d}" + (" --fastidious" if fast else "") + f", {label}"
    line = {
        "metric": metric, "value": n * args.steps / dt, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": wl, "seed": args.seed,
                   "kernels": {"enum_mode": args.enum_mode, "join_kernel": args.join_kernel, "cluster_kernel": args.cluster_kernel},
                   "l2": "inputs larger than L2 (packed database %.0f MB + %.0f MB tile store), no flush needed" % (n * (P_ + 12) / 1e6, 2 * n * (8 + P_) * 1.5 / 1e6),
                   "timing": "CUDA events on the engine's stream around the K steps, synchronize on both sides",
                   "wall_ms_per_step": 1e3 * wall / args.steps, "parallelism": "single GPU"},
        "phases_ms": {k: v["ms"] for k, v in roof["phases"].items()},
        "e2e": {"value": n * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * dt_e2e / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps,
                "api": "swb200_load_db_compact -> " + ("d1_index -> d1_network -> d1_cluster" + (" -> d1_fastidious" if fast else "") if d == 1 else "dn_cluster") + " (host arrays)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "parity": parity,
        "counters": {"links": m_links, **{k: cnt[k] for k in ("variants", "exact_compares", "rows_gathered", "cluster_rounds", "tile_overflow", "skew_fallbacks")}},
        "swarms": int((res["swarm_of"] == np.arange(n, dtype=np.uint32)).sum()),
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, cfg)
    emit(line)
    eng.close()


if __name__ == "__main__":
    main()
