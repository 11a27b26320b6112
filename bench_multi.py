"""bench.py, N > 1: ONE clustering job on N GPUs (torchrun, one rank per GPU), weak scaling — N x --amplicons amplicons.

Two layouts of the same job:
  replicated  (default for c2)  every GPU holds the packed database (4.2 GB at 80 M); each rank hashes its own n/N rows and ships
              16-byte records to the tile owners over NVLink; the join gathers rows locally
  sharded     (default for c5 = BASELINE configs[4])  a rank holds only its n/N rows; records carry their packed row to the tile
              owner, so no rank ever holds the whole database
In both the join tiles are sharded by hash range and the clustering by amplicon (block-cyclic), the exchanges are done by the
kernels themselves over peer memory; torch.distributed only sets the peer buffers up, does the barriers and the reductions of
the timings.  Parity (untimed): rank 0 clusters the gathered database with the single-GPU engine; every rank compares its rows.
"""
import json
import time

UNIT = "amplicons/s"


def build_weak_dataset(args, L, rank, world):
    """every rank generates and parses its own set (generator seed + rank: independent random centroids, no duplicate sequences
    in the union); the packed rows are gathered on every GPU and put into the reference's database order — abundance descending
    (src/db.cc:392-406; ties in rank, then header order) — with torch.  Setup only: not timed, not part of the product."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from bench import make_dataset
    from swarm_b200 import HostDb
    fa = f"/dev/shm/swb200_{args.amplicons}x{L}_s{args.seed + rank}.fa"
    make_dataset(args.amplicons, L, args.seed + rank, fa)
    db = HostDb(fa)
    stride = torch.tensor([db.stride, db.n], dtype=torch.int64, device="cuda")
    mx = stride.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    assert int(mx[1]) == db.n == args.amplicons, "every rank must hold the same number of amplicons"
    S = int(mx[0])
    w = torch.zeros((db.n, S), dtype=torch.int64, device="cuda")
    w[:, : db.stride] = torch.from_numpy(db.words.view(np.int64).reshape(db.n, db.stride)).cuda()
    ln = torch.from_numpy(db.len.view(np.int32)).cuda()
    ab = torch.from_numpy(db.abundance.view(np.int64)).cuda()
    db.close()
    W = torch.empty((world * args.amplicons, S), dtype=torch.int64, device="cuda")
    Ln = torch.empty(world * args.amplicons, dtype=torch.int32, device="cuda")
    Ab = torch.empty(world * args.amplicons, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(W, w)
    dist.all_gather_into_tensor(Ln, ln)
    dist.all_gather_into_tensor(Ab, ab)
    del w, ln, ab
    order = torch.sort(Ab, descending=True, stable=True).indices
    W, Ln, Ab = W[order].contiguous(), Ln[order].contiguous(), Ab[order].contiguous()
    del order
    torch.cuda.synchronize()
    return W, Ln, Ab, S


def run_multi(args, cfg, rank, world, local, emit, ClockSampler, peak, peaks_found):
    import numpy as np
    import torch
    import torch.distributed as dist
    from swarm_b200 import Engine
    from swarm_b200.ffi import dist_row_ids
    from swarm_b200.multi import all_gather_db, engine_stream, setup_dist_clustering, shard_rows

    n_gpu, L, d, fast, label, _case, metric = cfg
    mode = args.multi if args.multi != "auto" else ("sharded" if args.config == "c5" else "replicated")
    sharded = mode == "sharded"
    P_ = 8 * ((L + 31) // 32)

    def pinned(count, dtype):
        return torch.empty(count, dtype=dtype, pin_memory=True).numpy()

    W, Ln, Ab, stride = build_weak_dataset(args, L, rank, world)
    n = W.shape[0]
    first, count = shard_rows(n, rank, world)
    runs = torch.cat([torch.tensor([0], device="cuda"), torch.nonzero(Ab[1:] != Ab[:-1]).flatten() + 1,
                      torch.tensor([n], device="cuda")]).to(torch.int32).cpu().numpy().view(np.uint32)
    lmin, lmax = int(Ln.min()), int(Ln.max())
    # this rank's rows of the sorted database, in pinned host memory: what its host process hands to the C ABI
    pw, pl, pa = pinned(count * stride, torch.int64).view(np.uint64), pinned(count, torch.int32).view(np.uint32), pinned(count, torch.int64).view(np.uint64)
    pw[:] = W[first:first + count].reshape(-1).cpu().numpy().view(np.uint64)
    pl[:] = Ln[first:first + count].cpu().numpy().view(np.uint32)
    pa[:] = Ab[first:first + count].cpu().numpy().view(np.uint64)

    # compact form of the same rows for the replicated layout's upload: u16 lengths, and the abundance RUNS of the whole database
    # instead of an abundance array (expanded on every device)
    pl16 = pinned(count, torch.int16).view(np.uint16)
    pl16[:] = pl
    run_ab_t = Ab[torch.from_numpy(runs[:-1].astype(np.int64)).cuda()].cpu().numpy().view(np.uint64)
    prab, prst = pinned(run_ab_t.shape[0], torch.int64).view(np.uint64), pinned(runs.shape[0], torch.int32).view(np.uint32)
    prab[:], prst[:] = run_ab_t, runs
    eng = Engine(local, enum_mode=args.enum_mode, join_kernel=args.join_kernel, collect_stats=0, shard_rank=rank, shard_world=world,
                 tile_rows=1 if sharded else 0, job_min_len=lmin, job_max_len=lmax)
    if sharded:
        eng.load_db_rows(pw, stride, pl, pa, n, first, runs)
    else:
        eng.load_db_device(W.data_ptr(), stride, Ln.data_ptr(), Ab.data_ptr(), n)
    if rank != 0 or args.no_parity:
        del W, Ln, Ab                      # rank 0 keeps the gathered database for the parity leg
        torch.cuda.empty_cache()
    own_ids = dist_row_ids(n, rank, world)
    res = {k: pinned(own_ids.shape[0], torch.int32).view(np.uint32) for k in ("swarm_of", "generation", "parent")}
    h2d = (pw.nbytes + pl.nbytes + pa.nbytes + runs.nbytes) if sharded else (pw.nbytes + pl16.nbytes + prab.nbytes + prst.nbytes)
    d2h = 3 * 4 * own_ids.shape[0]
    ext = engine_stream(eng)
    inbox_bytes = setup_dist_clustering(eng, n)

    def device_step():
        eng.d1_index()
        eng.d1_network()
        eng.d1_cluster_dist(None)

    def e2e_step():
        if sharded:
            eng.load_db_rows(pw, stride, pl, pa, n, first, runs)
        else:
            eng.load_db_shard_compact(pw, stride, pl16, n, first, prab, prst)
            all_gather_db(eng, n, stride, with_abundance=False)
        eng.d1_index()
        eng.d1_network()
        return eng.d1_cluster_dist(res)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def timed(step, k):
        """K steps bracketed by barrier + synchronize on both sides; device time from CUDA events on the engine's stream"""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        t0 = time.perf_counter()
        ev0.record(ext)
        for _ in range(k):
            step()
        ev1.record(ext)
        sync_all()
        return ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - t0

    for _ in range(args.warmup):
        device_step()
    launches0 = eng.stats()["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    phase = {1: [], 2: [], 3: []}

    def device_step_logged():
        device_step()
        for p in phase:
            phase[p].append(eng.phase_seconds(p))

    dt, wall = timed(device_step_logged, args.steps)
    launches = eng.stats()["launches"] - launches0
    eng.set_option("collect_stats", 1)
    device_step()
    st = eng.stats()
    eng.set_option("collect_stats", 0)
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    dt_e2e, wall_e2e = timed(e2e_step, args.steps)
    clocks = sampler.finish() if rank == 0 else None
    free_b, total_b = torch.cuda.mem_get_info()

    # ---- parity (untimed): the single-GPU engine on the gathered database, rank 0; every rank compares its own rows
    parity = None
    if not args.no_parity:
        full = {k: torch.empty(n, dtype=torch.int32, device="cuda") for k in ("swarm_of", "generation", "parent")}
        if rank == 0:
            one = Engine(local, collect_stats=0)
            one.load_db_device(W.data_ptr(), stride, Ln.data_ptr(), Ab.data_ptr(), n)
            del W, Ln, Ab
            one.d1_index()
            one.d1_network()
            sw, gen, par = one.d1_cluster()
            one.close()
            for k, a in zip(("swarm_of", "generation", "parent"), (sw, gen, par)):
                full[k].copy_(torch.from_numpy(a.view(np.int32)))
            del sw, gen, par
        ok = 1
        for k in full:
            dist.broadcast(full[k], src=0)
            want = full[k].cpu().numpy().view(np.uint32)[own_ids]
            ok &= int(np.array_equal(want, res[k]))
        n_swarms_ref = int((full["swarm_of"].cpu().numpy().view(np.uint32) == np.arange(n, dtype=np.uint32)).sum())
        del full
        okt = torch.tensor([ok], dtype=torch.int64, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        parity = {"ok": bool(int(okt[0])), "vs": f"the single-GPU engine (hash-pinned to the reference at 10 M, see the N=1 line) run by rank 0 on the gathered "
                                                f"database of {n} amplicons; swarm / generation / parent of every rank's own rows compared on that rank",
                  "swarms": n_swarms_ref, "from": "the host arrays of the last end-to-end step"}

    times = torch.tensor([dt, dt_e2e, wall, wall_e2e] + [sum(phase[p]) / len(phase[p]) for p in (1, 2, 3)], dtype=torch.float64, device="cuda")
    stat_t = torch.tensor([st["variants"], st["exact_compares"], st["links"], st["rows_gathered"], int((res["swarm_of"] == own_ids).sum()),
                           st["tile_overflow"]], dtype=torch.int64, device="cuda")
    memt = torch.tensor([total_b - free_b], dtype=torch.int64, device="cuda")
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dist.all_reduce(stat_t, op=dist.ReduceOp.SUM)
    dist.all_reduce(memt, op=dist.ReduceOp.MAX)
    dt, dt_e2e, wall, wall_e2e, idx_s, net_s, clu_s = (float(x) for x in times)
    cnt = [int(x) for x in stat_t]

    if rank == 0:
        from bench import roofline_block
        per = n / world
        e_ = cnt[2] / world
        rec = (8 + P_) if sharded else 16
        ph = {
            "index": {"kernel": "k_ts_route + k_ts_scatter_inbox", "s": idx_s, "bytes": per * (P_ + 12) + 2 * per * rec * 2 + 2 * per * (8 + (P_ if sharded else 0)),
                      "formula": "per rank: own rows read, two records written to the owners' inboxes (NVLink) and read back, two tile records written"},
            "network": {"kernel": "k_ts_join", "s": net_s, "bytes": 2 * per * (8 + P_) + 8 * e_,
                        "formula": "per rank: every tile record + its packed row read once, links written"},
            "cluster": {"kernel": "k_cluster_bucket", "s": clu_s, "bytes": per * 28 + e_ * 36 + e_ * 16 * 2.6 * 2 * (world - 1) / world,
                        "formula": "per rank: relaxation words initialised and unpacked (28 B per owned amplicon), links routed, bucketed and walked (36 B per link), "
                                   "16-byte offers written to and read from the owners' logs (2.6 per link measured, the remote share)",
                        "rounds": st["cluster_rounds"]},
        }
        roof = roofline_block(peak, peaks_found, ph, None)
        wl = f"{n} x {L} bp synthetic amplicons, d=1 — ONE job, {args.amplicons} amplicons per GPU" + \
             (", BASELINE configs[4]" if (args.config == "c5" and n == 100_000_000) else " (BASELINE configs[4] shape)")
        emit({
            "metric": metric, "value": n * args.steps / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": wl, "seed": args.seed, "layout": mode,
                       "l2": "inputs larger than L2 (per GPU: %.0f MB of packed rows + %.0f MB tile store), no flush needed" % (per * (P_ + 12) / 1e6, 3 * per * (8 + (P_ if sharded else 0)) / 1e6),
                       "timing": "CUDA events on the engine's stream around the K steps, barrier + synchronize on both sides, max over ranks",
                       "wall_ms_per_step": 1e3 * wall / args.steps,
                       "parallelism": (f"packed database SHARDED by rows over {world} GPUs (no rank holds it all); " if sharded else f"packed database replicated on {world} GPUs; ") +
                                      "every rank hashes its own rows and routes the records to the tile owners over NVLink peer memory inside the kernel; "
                                      "join tiles sharded by hash range; clustering sharded by amplicon, links and label updates exchanged by the kernel over NVLink peer memory",
                       "hbm_used_bytes_max_rank": int(memt[0]), "peer_buffer_bytes_per_rank": int(inbox_bytes)},
            "phases_ms": {"index": 1e3 * idx_s, "network": 1e3 * net_s, "cluster": 1e3 * clu_s},
            "e2e": {"value": n * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": 1e3 * dt_e2e / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps,
                    "api": ("swb200_load_db_rows" if sharded else "swb200_load_db_shard_compact + all-gather of words and lengths") + " -> d1_index -> d1_network -> d1_cluster_dist (host arrays)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "parity": parity,
            "counters": {"records": cnt[0], "exact_compares": cnt[1], "links": cnt[2], "rows_gathered": cnt[3], "tile_overflow": cnt[5]},
            "swarms": cnt[4],
        })
    eng.close()
    dist.destroy_process_group()
