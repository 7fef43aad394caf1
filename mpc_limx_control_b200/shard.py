"""Instance sharding across GPUs/ranks: contiguous block partition by
instance, no collective on the solve path, final gather of the per-rank results.  Works with any
torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def partition(B, world, rank):
    """Contiguous block of instances owned by `rank`: (start, count). The first B % world ranks
    own one extra instance, so counts differ by at most one and cover [0, B) exactly."""
    base, extra = divmod(int(B), int(world))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def gather_rows(local, B, group=None, device=None):
    """Gather per-rank row blocks (numpy [count, ...]) into the full [B, ...] array on rank 0
    (None elsewhere).  Pads to the largest block so that a plain all_gather works on every backend."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return np.asarray(local)
    counts = [partition(B, world, r)[1] for r in range(world)]
    cmax = max(counts)
    t = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=torch.from_numpy(np.asarray(local)).dtype, device=device)
    t[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(t.device)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t, group=group)
    if rank != 0:
        return None
    return np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)
