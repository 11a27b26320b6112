"""Multi-GPU plumbing: one process per GPU, seeds sharded across ranks, ONE exchange
step — the all-gather(v) of the directed link lists — then replicated clustering.  torch.distributed is
only the transport (NCCL over NVLink on GPUs, gloo in the CPU tests); no algorithm lives here."""
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


def exchange_engine_links(eng, group=None):
    """the one data-path collective: gather every rank's link list and hand the union back to the engine"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ptr, m = eng.d1_links_device()
    if m:
        local = torch.as_tensor(_DevView(ptr, m * 8), device="cuda").reshape(-1, 2)
    else:
        local = torch.zeros((0, 2), dtype=torch.int32, device="cuda")
    merged = all_gather_links(local, group).contiguous()
    torch.cuda.synchronize()
    eng.d1_import_links_device(merged.data_ptr(), merged.shape[0])
