import sys, threading
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import helpers
from swarm_b200 import Engine, HostDb
from swarm_b200.ffi import compact_form, dist_buffer_bytes, dist_row_ids
fa = "/dev/shm/dbg_120k.fa"
helpers.make_fasta(fa, 120000, 150, 33, 0)
db = HostDb(fa)
one = Engine(0); one.load(db); one.d1_index(); one.d1_network(); ref_links = one.d1_export_links(); sw, gen, par = one.d1_cluster(); one.close()
ref_links = ref_links[np.lexsort((ref_links[:, 1], ref_links[:, 0]))]
n = db.n
def run(world, sharded, exchange, cluster, div=1):
    nbytes = dist_buffer_bytes(n, world, 4)
    bufs = [torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device="cuda") for _ in range(world)]
    ptrs = [b.data_ptr() for b in bufs]
    engs = [Engine(0, tile_rows=1 if sharded else 0, dist_grid_div=world * div, index_exchange=exchange, shard_rank=r, shard_world=world) for r in range(world)]
    _a, _b, rst = compact_form(db.len, db.abundance)
    per = (n + world - 1) // world
    for r, e in enumerate(engs):
        if sharded:
            lo, hi = min(per * r, n), min(per * (r + 1), n)
            e.load_db_rows(np.ascontiguousarray(db.words[lo * db.stride:hi * db.stride]), db.stride, np.ascontiguousarray(db.len[lo:hi]),
                           np.ascontiguousarray(db.abundance[lo:hi]), n, lo, rst)
        else:
            e.load(db)
        e.dist_setup(r, world, ptrs, nbytes)
        e.d1_reserve()
    torch.cuda.synchronize()
    outs, errs = [None] * world, []
    def work(r):
        try:
            engs[r].d1_index(); engs[r].d1_network()
            links = engs[r].d1_export_links()
            out = None
            if cluster:
                ids = dist_row_ids(n, r, world)
                out = {k: np.empty(ids.shape[0], dtype=np.uint32) for k in ("swarm_of", "generation", "parent")}
                engs[r].d1_cluster_dist(out)
                out = (ids, out)
            outs[r] = (links, out)
        except Exception as exc:
            errs.append((r, repr(exc)[:300]))
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    ok_links = ok_cl = None
    if not errs:
        links = np.concatenate([o[0] for o in outs]); links = links[np.lexsort((links[:, 1], links[:, 0]))]
        ok_links = bool(np.array_equal(links, ref_links))
        if cluster:
            ok_cl = all(np.array_equal(o[1][1]["swarm_of"], sw[o[1][0]]) and np.array_equal(o[1][1]["generation"], gen[o[1][0]]) and np.array_equal(o[1][1]["parent"], par[o[1][0]]) for o in outs)
    print(f"div={div} world={world} sharded={sharded} exchange={exchange} cluster={cluster}: errs={errs} links_ok={ok_links} cluster_ok={ok_cl}", flush=True)
    for e in engs:
        try: e.close()
        except Exception: pass
for args in ((2, True, 1, True, 1), (2, False, 1, True, 1), (4, True, 1, True, 1), (8, False, 1, True, 1), (3, False, 0, True, 1)):
    try:
        run(*args)
    except Exception as exc:
        print("FAILED", args, repr(exc)[:300], flush=True)
