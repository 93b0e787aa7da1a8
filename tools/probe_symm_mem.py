"""Probe: does torch.distributed._symmetric_memory give peer-mapped buffers on this box?
torchrun --nproc-per-node 2 tools/probe_symm_mem.py"""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{rank}")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok: world", hdl.world_size, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    print(rank, "peer value", float(peer[12345]), "(expect", float((rank + 1) % world + 1), ")", flush=True)
    peer[: 16] = 100.0 + rank  # a store into the peer's memory
    hdl.barrier()
    print(rank, "own head after peer store", t[:2].tolist(), flush=True)
    # bandwidth of a peer copy
    src = torch.empty(64 << 20, dtype=torch.uint8, device=f"cuda:{rank}")
    big = symm_mem.empty(64 << 20, dtype=torch.uint8, device=f"cuda:{rank}")
    h2 = symm_mem.rendezvous(big, dist.group.WORLD)
    pb = h2.get_buffer((rank + 1) % world, (64 << 20,), torch.uint8)
    torch.cuda.synchronize(); h2.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pb.copy_(src)
    e1.record(); torch.cuda.synchronize()
    print(rank, f"peer copy {10 * 64 / 1024 / (e0.elapsed_time(e1) / 1e3):.1f} GB/s", flush=True)
except Exception as ex:  # noqa: BLE001
    print(rank, "symmetric memory unavailable:", type(ex).__name__, str(ex)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
