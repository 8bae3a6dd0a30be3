"""Multi-GPU protocol behind the C ABI (csrc/comm.cu): packed 64-byte records, ONE all-gather,
device merge -- against the numpy statement of the same protocol (model_matching_b200/sharding.py)
and against the oracle's per-hypothesis results.  Multi-device cases need >= 2 GPUs and are skipped
on a single-GPU box (the round-end bench asserts the merged result on every multi-GPU run)."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle
from model_matching_b200 import sharding, synth

pytestmark = pytest.mark.gpu
K = 32


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def case(small_scene):
    sc, mpos, mnrm = small_scene
    T, _ = synth.make_hypotheses(3001, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=5, near_fraction=0.05)
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    olcp, oinl = est.score(T, threads=os.cpu_count() or 1)
    return sc, mpos, mnrm, T, olcp, oinl, est


def _same(a, b):
    return a.tobytes() == b.tobytes()


def test_single_rank_records_match_oracle(gpu_ctx, case):
    sc, mpos, mnrm, T, olcp, oinl, _ = case
    gpu_ctx.upload_model(mpos, mnrm)
    gpu_ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    for off in (0, 123456789012):
        rec, lcp, inl = gpu_ctx.score_sharded(T, off, K, want_local=True)
        assert np.array_equal(inl, oinl) and np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32))
        assert _same(rec, sharding.local_records(olcp, oinl, T, off, K))
    # fewer winners than K, and none at all
    far = T.copy(); far[:, 12:15] += 50.0
    far[7] = T[np.argmax(olcp)]
    rec = gpu_ctx.score_sharded(far, 10, 8)
    assert rec["index"].tolist() == [17] + [-1] * 7 and rec["lcp"][0] == olcp.max()
    rec = gpu_ctx.score_sharded(far[:5], 0, 4)
    assert (rec["index"] == -1).all() and (rec["lcp"] == 0).all() and (rec["T"] == 0).all()


def test_device_entry_point_and_pageable_path(gpu_ctx, case):
    import torch
    sc, mpos, mnrm, T, olcp, oinl, _ = case
    gpu_ctx.upload_model(mpos, mnrm)
    gpu_ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    want = sharding.local_records(olcp, oinl, T, 77, K)
    dT = torch.from_numpy(T).cuda()
    drec = torch.zeros(K * 64, dtype=torch.uint8, device="cuda")
    gpu_ctx.score_sharded_device(dT.data_ptr(), len(T), 77, K, drec.data_ptr())
    torch.cuda.synchronize()
    assert _same(drec.cpu().numpy().view(sharding.RECORD), want)
    n, mean, mx = gpu_ctx.kernel_ms_stats(reset=True)
    assert n >= 1 and 0 < mean <= mx
    assert gpu_ctx.kernel_ms_stats()[0] == 0
    # pinned (zero-copy) and staged host paths give the same records
    hT = torch.from_numpy(T).pin_memory()
    out = np.zeros(K, sharding.RECORD)
    gpu_ctx.score_sharded_ptr(hT.data_ptr(), len(T), 77, K, out)
    assert _same(out, want)
    os.environ["STOCS_NO_ZERO_COPY"] = "1"
    try:
        out2 = np.zeros(K, sharding.RECORD)
        gpu_ctx.score_sharded_ptr(hT.data_ptr(), len(T), 77, K, out2)
    finally:
        del os.environ["STOCS_NO_ZERO_COPY"]
    assert _same(out2, want)


def test_score_counters_are_exact(gpu_ctx, case):
    import torch
    sc, mpos, mnrm, T, olcp, oinl, est = case
    gpu_ctx.upload_model(mpos, mnrm)
    gpu_ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    dT = torch.from_numpy(T).cuda()
    c = gpu_ctx.score_counters(dT.data_ptr(), len(T))
    oc = est.score_counters(T, threads=os.cpu_count() or 1)
    assert c["queries"] == oc["queries"] == len(T) * len(mpos) and c["hypotheses"] == len(T)
    assert c["inliers"] == oc["inliers"] == int(oinl.sum())
    assert c["hits"] == oc["hits"]            # queries with a scene point within the threshold
    assert c["queries"] >= c["coarse_survivors"] >= c["brick_records"] >= c["queued"] >= c["hits"] >= c["inliers"]
    assert c["candidates"] >= c["queued"] and c["drains"] >= 1
    assert gpu_ctx.score_counters(dT.data_ptr(), len(T)) == c     # deterministic, no state carried over


def test_group_one_device_and_empty_list(case):
    from model_matching_b200 import Group
    sc, mpos, mnrm, T, olcp, oinl, _ = case
    g = Group([0])
    g.upload_model(mpos, mnrm)
    g.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    rec, lcp, inl = g.score_best(T, K, want_all=True)
    assert np.array_equal(inl, oinl) and np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32))
    assert _same(rec, sharding.local_records(olcp, oinl, T, 0, K))
    rec = g.score_best(T[:0], 4)
    assert (rec["index"] == -1).all()
    g.close()


@pytest.mark.skipif("_ngpu() < 2")
@pytest.mark.parametrize("ndev,H", [(2, 3001), (2, 1), (2, 33)])
def test_group_several_devices_equals_single(case, ndev, H):
    from model_matching_b200 import Group
    sc, mpos, mnrm, T, olcp, oinl, _ = case
    g = Group(list(range(min(ndev, _ngpu()))))
    g.upload_model(mpos, mnrm)
    g.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    rec, lcp, inl = g.score_best(T[:H], K, want_all=True)
    assert np.array_equal(inl, oinl[:H]) and np.array_equal(lcp.view(np.uint32), olcp[:H].view(np.uint32))
    assert _same(rec, sharding.local_records(olcp[:H], oinl[:H], T[:H], 0, K))
    g.close()


def _rank_main(rank, world, idfile, q):
    import model_matching_b200 as mm
    sc = synth.make_scene(n_points=40000, extent=(0.7, 0.5, 0.5), n_objects=5, seed=7)
    mpos, mnrm = synth.make_model(256)
    T, _ = synth.make_hypotheses(3001, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=5, near_fraction=0.05)
    ctx = mm.Context(rank)
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(mm.comm_unique_id())
        os.rename(idfile + ".tmp", idfile)
    import time
    while not os.path.exists(idfile):
        time.sleep(0.01)
    ctx.comm_init(open(idfile, "rb").read(), rank, world)
    lo, hi = mm.shard_range(len(T), rank, world)
    rec = ctx.score_sharded(T[lo:hi], lo, K)
    q.put((rank, rec.tobytes()))
    ctx.close()


@pytest.mark.skipif("_ngpu() < 2")
@pytest.mark.timeout(300)
def test_two_processes_one_allgather(case, tmp_path):
    import torch.multiprocessing as mp
    sc, mpos, mnrm, T, olcp, oinl, _ = case
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_rank_main, args=(r, 2, str(tmp_path / "nccl_id"), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sharding.local_records(olcp, oinl, T, 0, K).tobytes()
    assert got[0] == want and got[1] == want
