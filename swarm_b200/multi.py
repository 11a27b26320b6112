"""Multi-GPU plumbing for the Python hosts (tests, bench.py): one process per GPU.  torch.distributed only sets up the
peer-visible buffers (symmetric memory), optionally all-gathers a replicated database, and reduces timings — the exchanges of the
job itself (index records, links, label offers) are done by the engine's kernels over NVLink peer memory (csrc/d1_tsroute.cuh,
d1_bucket.cuh).  The C++ host does the same without torch: swb200_dist_setup_local (one process, one thread per GPU; host/main.cc).
The link all-gather(v) + replicated clustering of round 1 is kept as a cross-check (exchange_engine_links)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, batch: int, rank: int, world: int):
    """the engine's seed range for (rank, world): contiguous, batch-aligned (engine.cu: run_network)"""
    n_padded = (n + batch - 1) // batch * batch
    nb = n_padded // batch
    per = (nb + world - 1) // world
    b0 = min(per * rank, nb)
    b1 = min(b0 + per, nb)
    lo, hi = b0 * batch, min(b1 * batch, n)
    return (min(lo, hi), hi)


def all_gather_links(pairs: torch.Tensor, group=None) -> torch.Tensor:
    """all-gatherv of (m_r, 2) int32/uint32 link tensors: counts first, then one padded all_gather_into_tensor,
    then compaction.  Works for CUDA tensors (NCCL) and CPU tensors (gloo)."""
    world = dist.get_world_size(group)
    if world == 1:
        return pairs
    dev = pairs.device
    flat = pairs.reshape(-1).view(torch.int32) if pairs.dtype != torch.int32 else pairs.reshape(-1)
    cnt = torch.tensor([flat.numel()], dtype=torch.int64, device=dev)
    cnts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    cnts = cnts.tolist()
    mx = max(max(cnts), 1)
    mine = torch.zeros(mx, dtype=torch.int32, device=dev)
    mine[: flat.numel()] = flat
    allb = torch.empty(world * mx, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allb, mine, group=group)
    merged = torch.cat([allb[r * mx: r * mx + cnts[r]] for r in range(world)])
    return merged.reshape(-1, 2)


class _DevView:
    """expose a raw CUDA pointer to torch through __cuda_array_interface__ (no copy)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def device_view(ptr: int, nbytes: int) -> torch.Tensor:
    """int32 torch view of `nbytes` of device memory at `ptr` (no copy, no ownership)"""
    return torch.as_tensor(_DevView(ptr, nbytes), device="cuda")


def shard_rows(n_total: int, rank: int, world: int):
    """(first, count) of the database rows rank `rank` uploads: equal shards of ceil(n/world), the last one short"""
    per = (n_total + world - 1) // world
    first = min(per * rank, n_total)
    return first, min(per, n_total - first)


def engine_stream(eng):
    """the engine's CUDA stream as a torch stream: collectives issued under it are ordered after the engine's
    kernels and before its next ones without a host synchronisation"""
    return torch.cuda.ExternalStream(eng.stream())


def all_gather_db(eng, n_total: int, stride: int, group=None, with_abundance: bool = True):
    """the one exchange of the upload path: every rank pushed its own rows over PCIe (swb200_load_db_shard); the
    rows of the other ranks arrive device-to-device over NVLink (NCCL all-gather, in place on the engine's buffers)"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        eng.db_commit()
        return
    per = (n_total + world - 1) // world
    w_ptr, l_ptr, a_ptr = eng.db_device()
    with torch.cuda.stream(engine_stream(eng)):
        # (after swb200_load_db_shard_compact every device already holds all abundances: expanded from the run table)
        for ptr, row_bytes in ((w_ptr, stride * 8), (l_ptr, 4)) + (((a_ptr, 8),) if with_abundance else ()):
            whole = device_view(ptr, per * world * row_bytes)
            each = per * row_bytes // 4
            # in place: this rank's shard already lies at its slot of the output (NCCL's in-place all-gather layout), no staging copy
            dist.all_gather_into_tensor(whole, whole[rank * each:(rank + 1) * each], group=group)
    eng.db_commit()


def exchange_engine_links(eng, group=None):
    """the one data-path collective: gather every rank's link list and hand the union back to the engine"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ptr, m = eng.d1_links_device()
    with torch.cuda.stream(engine_stream(eng)):
        if m:
            local = device_view(ptr, m * 8).reshape(-1, 2)
        else:
            local = torch.zeros((0, 2), dtype=torch.int32, device="cuda")
        merged = all_gather_links(local, group).contiguous()
        eng.d1_import_links_device(merged.data_ptr(), merged.shape[0])      # stream-ordered D2D copy + sync inside


def setup_dist_clustering(eng, n_total: int, items_per_amplicon: int = 4, group=None):
    """peer-visible inboxes for swb200_d1_cluster_dist: one torch symmetric-memory buffer per rank (CUDA VMM / fabric
    handles exchanged by torch), its peer addresses handed to the engine as plain pointers.  Collective."""
    import torch.distributed._symmetric_memory as symm
    from .ffi import dist_buffer_bytes
    group = group or dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nbytes = dist_buffer_bytes(n_total, world, items_per_amplicon)
    buf = symm.empty((nbytes + 3) // 4, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
    hdl = symm.rendezvous(buf, group)
    buf.zero_()
    torch.cuda.synchronize()
    eng.dist_setup(rank, world, [int(p) for p in hdl.buffer_ptrs], nbytes)
    torch.cuda.synchronize()
    dist.barrier(group)
    eng._dist_keep = (buf, hdl)          # the engine only borrows the memory
    return nbytes
