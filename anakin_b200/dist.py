"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by independent
images, so the only collective is a ONE-TIME broadcast of the model bytes from the rank that built
/ loaded them (NCCL over NVLink on the GPU box, gloo in the CPU tests); the per-step path has no
collective at all.  The reference has no multi-GPU support: its Worker keeps
every replica on one device (framework/core/net/worker.cpp:10-53).
"""
import numpy as np


def broadcast_bytes(blob, src=0, device="cpu"):
    """Broadcast a bytes object from rank `src` to every rank; returns the bytes on all ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return blob
    rank = dist.get_rank()
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    if rank == src:
        buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy())


def broadcast_weight_arena(device, src=0):
    """One NCCL broadcast of the packed-weight arena of `device` (every conv / fc weight image, bias and scale table the
    Nets of this process hold, contiguous in creation order) from rank `src`. The other ranks must have initialised
    their Nets under api.weight_arena_set_receive(True): identical plans and buffers, nothing folded, quantised or
    packed on their hosts. Returns (bytes, milliseconds)."""
    import time
    import torch
    import torch.distributed as dist
    from . import api
    nbytes = api.weight_arena_flat_bytes(device)
    if not dist.is_initialized() or dist.get_world_size() == 1 or nbytes == 0:
        return nbytes, 0.0
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([nbytes], dtype=torch.int64, device="cuda"))
    if any(int(s.item()) != nbytes for s in sizes):
        raise RuntimeError("weight arenas differ between ranks: %s" % [int(s.item()) for s in sizes])
    flat = torch.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", device))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if dist.get_rank() == src:
        api.weight_arena_export(device, flat.data_ptr(), nbytes)
    dist.broadcast(flat, src)
    if dist.get_rank() != src:
        api.weight_arena_import(device, flat.data_ptr(), nbytes)
    torch.cuda.synchronize()
    return nbytes, (time.perf_counter() - t0) * 1e3


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` requests owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, device="cpu"):
    """all_gather of per-rank [n_i, C] float32 results into rank order (used by tests / examples
    when a single entry point wants every shard's logits; not on the benchmark path)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    t = torch.as_tensor(np.ascontiguousarray(local, np.float32), device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    pad[: t.shape[0]] = t
    outs = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(outs, pad)
    return np.concatenate([o[: int(s.item())].cpu().numpy() for o, s in zip(outs, sizes)], axis=0)
