"""N>1 protocol on CPU: hypothesis sharding, 64-byte records, ONE all-gather, deterministic merge --
world_size 2 over gloo.  The product runs the same protocol in C/CUDA over NCCL (csrc/comm.cu);
tests/test_comm_gpu.py compares that implementation with the numpy statement used here."""
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


def _inputs(H):
    rng = np.random.default_rng(42)
    lcp = rng.uniform(0, 1, H).astype(np.float32)
    lcp[rng.integers(0, H, 40)] = np.float32(0.999)      # ties on the maximum across shards
    lcp[rng.integers(0, H, H // 3)] = 0                  # hypotheses that match nothing
    inl = rng.integers(0, 512, H).astype(np.int32)
    T = rng.normal(size=(H, 16)).astype(np.float32)
    return lcp, inl, T


def _worker(rank, world, port, H, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lcp, inl, T = _inputs(H)
    lo, hi = sharding.shard_range(H, rank, world)
    local = sharding.local_records(lcp[lo:hi], inl[lo:hi], T[lo:hi], lo, K)
    gathered = sharding.all_gather_records(dist, local)          # the one collective
    merged = sharding.merge_records(gathered, K)
    q.put((rank, merged.tobytes()))
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


def test_shard_range_matches_the_library():
    import model_matching_b200 as mm
    for H in (0, 1, 7, 1000, 1000001, 10**7):
        for world in (1, 2, 3, 4, 8, 64):
            for r in range(world):
                assert mm.shard_range(H, r, world) == sharding.shard_range(H, r, world)


def test_record_layout_and_merge():
    import model_matching_b200 as mm
    assert mm.RECORD == sharding.RECORD and sharding.RECORD.itemsize == 64
    lcp, inl, T = _inputs(1000)
    whole = sharding.local_records(lcp, inl, T, 0, K)
    parts = [sharding.local_records(lcp[lo:hi], inl[lo:hi], T[lo:hi], lo, K)
             for lo, hi in (sharding.shard_range(1000, r, 3) for r in range(3))]
    merged = sharding.merge_records(np.concatenate(parts), K)
    assert merged.tobytes() == whole.tobytes()
    # the 12 floats are rows 0..2 of the column-major 4x4, row-major
    i = int(whole["index"][0])
    assert np.array_equal(whole["T"][0], T[i].reshape(4, 4).T[:3].reshape(12))
    # nothing scored: every slot empty
    z = sharding.merge_records(sharding.local_records(np.zeros(5, np.float32), np.zeros(5, np.int32), T[:5], 0, K), K)
    assert (z["index"] == -1).all() and (z["lcp"] == 0).all()


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_global_reduction():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    H = 10007
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=100)
        assert p.exitcode == 0
    lcp, inl, T = _inputs(H)
    whole = sharding.local_records(lcp, inl, T, 0, K)
    assert got[0] == got[1] == whole.tobytes()                   # every rank ends with the global top-K
    assert (int(whole["index"][0]), float(whole["lcp"][0])) == oracle.best(lcp)
