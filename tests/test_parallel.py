"""Host logic of the multi-GPU path, on CPU with the gloo backend and world_size 2."""
import importlib.util
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_parallel():
    spec = importlib.util.spec_from_file_location("ab200_parallel", os.path.join(ROOT, "release-1.0_b200", "parallel.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_shard_frontier_covers_everything_once():
    par = _load_parallel()
    for n in (0, 1, 31, 32, 33, 1000, 101520):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a, b = par.shard_frontier(n, r, world)
                assert 0 <= a <= b <= n and (a % 32 == 0 or a == n)
                seen.extend(range(a, b))
            assert seen == list(range(n))


def test_deal_blocks_partitions_the_frontier():
    par = _load_parallel()
    for n in (0, 1, 63, 64, 65, 1000, 101520):
        for world in (1, 2, 4, 8):
            parts = [par.deal_blocks(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if parts else np.zeros(0)
            assert np.array_equal(allidx, np.arange(n))
            # whole blocks stay together and the shares differ by at most one block
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 64
            for p in parts:
                if len(p):
                    assert np.all((np.diff(p) == 1) | (p[1:] % 64 == 0))


def _fake_lap(locs, boxes):
    """Stand-in for the device lap with the same contract: every cell reaches a few keys derived from its box."""
    b = np.asarray(boxes)[:, 0].astype(np.int64)
    reach = np.unique(np.concatenate([b * 3, b * 3 + 1, b * 3 + 2]).astype(np.uint64))
    evolve = np.unique((b * 3 + 2).astype(np.uint64))
    return reach, evolve


def _lap_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = _load_parallel()
    n = 1000
    locs = np.zeros((n,), dtype=np.int32)
    boxes = np.zeros((n, 10)); boxes[:, 0] = np.arange(n) // 2          # neighbouring cells reach overlapping keys
    reach, evolve = par.sharded_lap(_fake_lap, locs, boxes, dist)
    ref_reach, ref_evolve = _fake_lap(locs, boxes)
    q.put((rank, bool(np.array_equal(reach, ref_reach) and np.array_equal(evolve, ref_evolve)), int(reach.shape[0])))
    dist.destroy_process_group()


def test_sharded_lap_gloo_world2_equals_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = _load_parallel()
    rng = np.random.RandomState(1234)
    universe = np.unique(rng.randint(0, 1 << 40, size=5000).astype(np.uint64))
    a, b = par.shard_frontier(universe.shape[0], rank, world)
    # overlapping slices: neighbouring ranks reach some of the same cells
    mine = np.unique(np.concatenate([universe[a:b], universe[max(0, a - 50):a]]))
    union = par.exchange_keys(mine, dist)
    q.put((rank, bool(np.array_equal(union, universe)), int(union.shape[0])))
    dist.destroy_process_group()


def test_exchange_keys_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
