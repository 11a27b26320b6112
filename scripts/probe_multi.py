"""2-GPU probe: NCCL transport + all-gather bandwidth, peer access, torch symmetric memory (peer pointers)."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def log(*a):
    if rank == 0: print(*a, flush=True)
log("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
for mb in (1, 70, 520):
    x = torch.ones(mb * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")
    out = torch.empty(world * x.numel(), dtype=torch.int32, device="cuda")
    for _ in range(3): dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    log(f"all_gather {mb} MB/rank: {ms:.3f} ms  -> {mb * (world - 1) / ms:.1f} GB/s recv per rank")
t = torch.zeros(1, device="cuda")
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    dist.all_reduce(t); torch.cuda.synchronize()
log(f"all_reduce(1 elem)+sync latency: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us")
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty(1 << 20, dtype=torch.int32, device=torch.device("cuda", local))
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    log("symm_mem ok: buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
        "multicast", hdl.has_multicast_support(torch.device("cuda", local).type, local) if hasattr(hdl, "has_multicast_support") else None)
    buf.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.int32)
    log("peer read:", peer[:4].tolist())
    hdl.barrier()
except Exception as e:
    log("symm_mem FAILED:", repr(e))
dist.destroy_process_group()
