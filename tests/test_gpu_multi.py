"""Multi-GPU job on ONE GPU (-m gpu): W engine contexts on cuda:0 play the W ranks — their "peer" buffers are ordinary device
buffers of the same GPU, every rank runs on its own host thread and stream, and the persistent clustering kernels share the
SMs (option dist_grid_div).  This exercises exactly the code a multi-GPU job runs (index exchange through the inboxes,
sharded join, peer-memory clustering) where the driver has a single GPU; on real peers only the addresses differ.
Parity: every rank's rows == the single-GPU engine == the CPU oracle.  (At most 4 ranks: the ranks' streams must not share a
hardware queue — CUDA_DEVICE_MAX_CONNECTIONS is 8 — or a kernel queues behind a peer's kernel that is waiting for it.)"""
import threading

import numpy as np
import pytest

import helpers
from helpers import Oracle
from swarm_b200 import Engine, HostDb
from swarm_b200.ffi import compact_form, dist_buffer_bytes, dist_row_ids

pytestmark = pytest.mark.gpu


def run_virtual(db, world, sharded, steps=2, **opt):
    import torch
    n = db.n
    nbytes = dist_buffer_bytes(n, world, 4)
    bufs = [torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device="cuda") for _ in range(world)]
    ptrs = [b.data_ptr() for b in bufs]
    engs = [Engine(0, tile_rows=1 if sharded else 0, dist_grid_div=world, collect_stats=1, **dict({"index_exchange": 2}, **opt)) for _ in range(world)]
    _l16, _rab, rst = compact_form(db.len, db.abundance)
    per = (n + world - 1) // world
    for r, e in enumerate(engs):
        e.set_option("job_min_len", int(db.len.min()))
        e.set_option("job_max_len", int(db.len.max()))
        if sharded:
            lo, hi = min(per * r, n), min(per * (r + 1), n)
            e.load_db_rows(np.ascontiguousarray(db.words[lo * db.stride:hi * db.stride]), db.stride, np.ascontiguousarray(db.len[lo:hi]),
                           np.ascontiguousarray(db.abundance[lo:hi]), n, lo, rst)
        else:
            e.load(db)
        e.dist_setup(r, world, ptrs, nbytes)
        e.d1_reserve()
    torch.cuda.synchronize()
    outs = [None] * world
    errs = []

    def work(r):
        try:
            for _ in range(steps):                          # twice: inboxes, epochs and counters must be reusable
                ids = dist_row_ids(n, r, world)
                out = {k: np.empty(ids.shape[0], dtype=np.uint32) for k in ("swarm_of", "generation", "parent")}
                engs[r].d1_index()
                engs[r].d1_network()
                engs[r].d1_cluster_dist(out)
                outs[r] = (ids, out, engs[r].d1_export_links(), engs[r].stats())
        except Exception as exc:                            # noqa: BLE001
            errs.append((r, repr(exc)))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in engs:
        e.close()
    assert not errs, errs
    return outs


@pytest.mark.parametrize("world,sharded", [(2, True), (2, False), (4, True), (3, False), (3, True)])
def test_virtual_ranks_vs_oracle(built, tmp_path, world, sharded):
    fa = helpers.make_fasta(tmp_path / "s.fa", 120000, 150, 31 + world, 0)
    db = HostDb(fa)
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    outs = run_virtual(db, world, sharded)
    links = np.concatenate([o[2] for o in outs])
    links = links[np.lexsort((links[:, 1], links[:, 0]))]
    assert np.array_equal(links, orc.links()), "union of the ranks' links differs from the oracle"
    assert sum(o[3]["variants"] for o in outs) == 2 * db.n          # every record reached exactly one tile owner
    for ids, out, _l, _s in outs:
        assert np.array_equal(out["swarm_of"], orc.swarm_of[ids])
        assert np.array_equal(out["generation"], orc.generation[ids])
        assert np.array_equal(out["parent"], orc.parent[ids])


def test_virtual_ranks_tie_heavy_and_mixed_lengths(built, tmp_path):
    parts = []
    for i, L in enumerate((40, 150, 151, 260)):
        p = helpers.make_fasta(tmp_path / f"p{i}.fa", 20000, L, 70 + i, 1)
        parts.append(open(p, "rb").read().replace(b">s", b">l%d_" % L))
    db = HostDb(text=b"".join(parts))
    orc = Oracle(db)
    orc.network()
    orc.cluster()
    # cluster_gen_bits = 2: a swarm deeper than the packed relaxation word holds -> every rank goes again with 32-bit generations
    for world, sharded, opt in ((2, True, {}), (4, False, {}), (2, False, {"dist_kernel": 1}), (2, False, {"index_exchange": 0}),
                                (3, True, {"cluster_gen_bits": 2}), (2, False, {"cluster_pack": 0})):
        outs = run_virtual(db, world, sharded, steps=1, **opt)
        assert all(o[3]["cluster_unpacked_reruns"] == (1 if "cluster_gen_bits" in opt else 0) for o in outs), opt
        links = np.concatenate([o[2] for o in outs])
        links = links[np.lexsort((links[:, 1], links[:, 0]))]
        assert np.array_equal(links, orc.links())
        for ids, out, _l, _s in outs:
            assert np.array_equal(out["swarm_of"], orc.swarm_of[ids]) and np.array_equal(out["generation"], orc.generation[ids])
            assert np.array_equal(out["parent"], orc.parent[ids])
