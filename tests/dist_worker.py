"""torchrun worker (one rank per GPU): the multi-GPU path of ONE d=1 job — join tiles sharded by hash range,
clustering sharded by amplicon range with the exchange over peer memory (swb200_d1_cluster_dist) — must return,
on every rank, exactly the rows a single-GPU run returns.  Also exercises the sharded upload (load_db_shard + NCCL
all-gather) and the replicated-clustering path (link all-gather).  usage: dist_worker.py <fasta> [repeat]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from swarm_b200 import Engine, HostDb  # noqa: E402
from swarm_b200.ffi import compact_form, dist_row_ids  # noqa: E402
from swarm_b200.multi import all_gather_db, exchange_engine_links, setup_dist_clustering, shard_rows  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    db = HostDb(sys.argv[1])
    repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n = db.n
    # single-GPU answer
    ref = Engine(local)
    ref.load(db)
    ref.d1_index()
    ref.d1_network()
    sw, gen, par = ref.d1_cluster()
    ref.close()
    first, count = shard_rows(n, rank, world)
    # sharded upload + sharded join
    eng = Engine(local, shard_rank=rank, shard_world=world)
    w = db.words.reshape(n, db.stride)[first:first + count].reshape(-1).copy()
    if os.environ.get("SWB200_TEST_PLAIN_SHARD"):
        eng.load_db_shard(w, db.stride, db.len[first:first + count].copy(), db.abundance[first:first + count].copy(), n, first)
        all_gather_db(eng, n, db.stride)
    else:                                                      # compact: u16 lengths, abundance runs of the whole database, no abundance exchange
        l16, rab, rst = compact_form(db.len, db.abundance)
        eng.load_db_shard_compact(w, db.stride, l16[first:first + count].copy(), n, first, rab, rst)
        all_gather_db(eng, n, db.stride, with_abundance=False)
    setup_dist_clustering(eng, n)
    ids = dist_row_ids(n, rank, world).astype(np.int64)       # the rows this rank owns (block-cyclic)
    for it in range(repeat):
        eng.d1_index()
        eng.d1_network()
        out = {k: np.empty(ids.shape[0], dtype=np.uint32) for k in ("swarm_of", "generation", "parent")}
        eng.d1_cluster_dist(out)
        for k, full in (("swarm_of", sw), ("generation", gen), ("parent", par)):
            assert np.array_equal(out[k], full[ids]), (rank, it, k, int((out[k] != full[ids]).sum()))
    # replicated clustering over the gathered links gives the same rows
    eng.d1_index()
    eng.d1_network()
    exchange_engine_links(eng)
    eng.d1_cluster(want=())
    out = {k: np.empty(count, dtype=np.uint32) for k in ("swarm_of", "generation", "parent")}
    eng.d1_get_cluster(first, count, out)
    for k, full in (("swarm_of", sw), ("generation", gen), ("parent", par)):
        assert np.array_equal(out[k], full[first:first + count]), (rank, "replicated", k)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} dist ok: n={n} rows [{first},{first + count})", flush=True)


if __name__ == "__main__":
    main()
