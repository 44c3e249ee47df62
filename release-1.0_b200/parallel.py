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


def deal_blocks(n_cells, rank, world, block=64):
    """Indices of the frontier cells of `rank` when blocks of `block` consecutive cells (one CUDA block of the evolution
    kernel) are dealt round-robin to the ranks. The frontier is enumerated in tree order, so consecutive cells are
    neighbours with similar step counts; dealing blocks instead of cutting contiguous slices spreads the long-running
    regions over all GPUs."""
    idx = np.arange(n_cells, dtype=np.int64)
    return idx[(idx // block) % world == rank]


def sharded_lap(lap, locs, boxes, dist=None, device=None, block=64):
    """One lap of the analyser on all ranks: `lap(locs, boxes) -> (reach_keys, evolve_keys)` (sorted unique uint64
    arrays, e.g. Evolver.upper_reach_evolve) runs on this rank's share of the frontier; the reached and the evolved
    keys are then exchanged, so every rank returns the sets of the whole lap (the reference adjoins the per-worker
    results under a mutex, include/workers.h:141-151)."""
    locs = np.asarray(locs); boxes = np.asarray(boxes)
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = deal_blocks(locs.shape[0], rank, world, block)
    reach, evolve = lap(np.ascontiguousarray(locs[mine]), np.ascontiguousarray(boxes[mine]))
    return exchange_keys(reach, dist, device), exchange_keys(evolve, dist, device)


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
