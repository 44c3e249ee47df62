"""Multi-GPU plumbing of the lap: frontier cells are independent units (the reference farms them to threads with no
communication, include/workers.h:110-153), so each rank evolves a contiguous, warp-aligned slice of the frontier and
the only exchange step is the all-gather of the newly reached cell keys at the end of the lap."""
import numpy as np


def shard_frontier(n_cells, rank, world, align=32):
    """[start, stop) of rank's slice: equal contiguous slices rounded up to a multiple of `align` cells (one warp)."""
    per = (n_cells + world - 1) // world
    per = (per + align - 1) // align * align
    start = min(n_cells, rank * per)
    stop = min(n_cells, start + per)
    return start, stop


def exchange_keys(keys, dist=None, device=None):
    """Union over ranks of sorted duplicate-free uint64 key arrays: all-gather of the counts, then of the padded
    payloads (NCCL over NVLink on GPUs, gloo in the CPU tests). Returns the sorted duplicate-free union."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(keys, dtype=np.uint64)
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    local = torch.from_numpy(np.ascontiguousarray(keys).view(np.int64)).to(dev)
    count = torch.tensor([local.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count)
    sizes = [int(c.item()) for c in counts]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=torch.int64, device=dev)
    pad[:local.numel()] = local
    parts = [torch.empty(mx, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(parts, pad)
    allk = torch.cat([p[:s] for p, s in zip(parts, sizes)]).cpu().numpy().view(np.uint64)
    return np.unique(allk)
