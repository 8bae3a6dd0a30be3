"""N>1 host logic on CPU: hypothesis sharding + the one all-gather + deterministic merge,
world_size 2 over gloo (the NCCL run uses exactly this code on CUDA tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from model_matching_b200 import sharding

K = 8


def _local_topk(lcp, lo, k):
    order = np.lexsort((np.arange(lcp.size), -lcp.astype(np.float64)))
    order = order[lcp[order] > 0][:k]
    idx = np.full(k, -1, np.int64); val = np.zeros(k, np.float32)
    idx[:order.size] = order + lo; val[:order.size] = lcp[order]
    return idx, val


def _worker(rank, world, port, H, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    lcp = rng.uniform(0, 1, H).astype(np.float32)
    lcp[rng.integers(0, H, 40)] = np.float32(0.999)      # ties on the maximum across shards
    lo, hi = sharding.shard_range(H, rank, world)
    li, lv = _local_topk(lcp[lo:hi], lo, K)
    gi, gv = sharding.all_gather_topk(dist, torch.from_numpy(li), torch.from_numpy(lv))
    mi, mv = sharding.merge_topk(gi.numpy(), gv.numpy(), K)
    if rank == 0:
        q.put((mi, mv, lcp))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_list():
    for H in (0, 1, 7, 1000, 1000001):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_merge_rule_matches_first_strict_maximum():
    idx = np.array([5, 2, -1, 9, 7], np.int64)
    val = np.array([0.5, 0.5, 0.0, 0.25, 0.0], np.float32)
    mi, mv = sharding.merge_topk(idx, val, 4)
    assert mi.tolist() == [2, 5, 9, -1] and mv.tolist() == [0.5, 0.5, 0.25, 0.0]
    assert sharding.best_of(np.array([-1, -1]), np.array([0.0, 0.0], np.float32)) == (-1, 0.0)


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_global_reduction():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    H = 10007
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, q)) for r in range(2)]
    for p in procs:
        p.start()
    mi, mv, lcp = q.get(timeout=100)
    for p in procs:
        p.join(timeout=100)
        assert p.exitcode == 0
    gi, gv = _local_topk(lcp, 0, K)
    assert np.array_equal(mi, gi) and np.array_equal(mv, gv)
    assert (int(mi[0]), float(mv[0])) == oracle.best(lcp)
