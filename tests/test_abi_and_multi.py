"""CPU tests: the C-ABI libraries load and export every symbol their headers declare; the N>1 host
logic (seed sharding + link all-gatherv) works over gloo with world_size 2."""
import ctypes
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared(header):
    txt = (ROOT / "include" / header).read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sw(?:b200|bh)_\w+)\s*\(", txt)))


@pytest.mark.parametrize("header,lib", [("swarm_b200.h", "libswarm_b200.so"), ("swarm_b200_host.h", "libswarm_b200_host.so")])
def test_library_exports_every_declared_symbol(built, header, lib):
    L = ctypes.CDLL(str(ROOT / "swarm_b200" / lib))
    names = declared(header)
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_engine_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from swarm_b200 import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)       # no CUDA device: the product has no CPU fallback


def test_shard_ranges_partition_the_seeds():
    from swarm_b200.multi import shard_range
    for n, batch in [(1000, 6), (10_000_000, 6), (7, 2), (1, 8), (333, 8)]:
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, batch, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0]
            assert all(lo % batch == 0 or lo == hi for lo, hi in rs)


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import helpers
from swarm_b200 import HostDb
from swarm_b200.multi import all_gather_links, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
db = HostDb(os.path.join(sys.argv[1], "tests", "golden", "c1_1k_150.fasta"))
orc = helpers.Oracle(db); orc.network()
links = orc.links()
lo, hi = shard_range(db.n, 6, rank, world)
mine = links[(links[:, 0] >= lo) & (links[:, 0] < hi)]        # what this rank's engine shard would find (FULL mode)
merged = all_gather_links(torch.from_numpy(mine.astype(np.int32)))
got = merged.numpy().astype(np.uint32)
got = got[np.lexsort((got[:, 1], got[:, 0]))]
assert np.array_equal(got, links), (rank, got.shape, links.shape)
empty = all_gather_links(torch.zeros((0, 2), dtype=torch.int32) if rank else torch.from_numpy(mine.astype(np.int32)))
assert empty.shape[0] == (links[(links[:, 0] < shard_range(db.n, 6, 0, world)[1])].shape[0])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("n", [1, 4095, 4096, 4097, 100000, 1234567])
@pytest.mark.parametrize("world", [1, 2, 3, 8, 16])
def test_dist_rows_partition_the_amplicons(built, n, world):
    """multi-GPU clustering returns, on rank r, the rows of the amplicons it owns (block-cyclic, blocks of 4 096 ids:
    swb200_dist_row_count / swb200_dist_row_id, pure host functions): over all ranks every amplicon exactly once, ascending on a
    rank, and the peer buffer size the host is told to allocate grows with the job"""
    from swarm_b200.ffi import dist_buffer_bytes, dist_row_ids, engine_lib
    L = engine_lib()
    seen = np.zeros(n, dtype=np.uint8)
    for r in range(world):
        ids = dist_row_ids(n, r, world).astype(np.int64)
        assert ids.shape[0] == L.swb200_dist_row_count(n, r, world)
        assert np.all(np.diff(ids) > 0) and (ids.shape[0] == 0 or ids[-1] < n)
        assert np.all((ids // 4096) % world == r)
        for i in (0, ids.shape[0] // 2, ids.shape[0] - 1):
            if 0 <= i < ids.shape[0]:
                assert L.swb200_dist_row_id(r, world, int(i)) == ids[i]
        seen[ids] += 1
    assert np.all(seen == 1)
    assert dist_buffer_bytes(2 * n, world) >= dist_buffer_bytes(n, world) > 4096


def test_link_allgather_world2_gloo(built, tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("ok" in o for o in outs)
