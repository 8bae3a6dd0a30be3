"""GPU parity of LCP scoring (reference src/stocs.cpp:1006-1041) against the CPU oracle:
inlier counts AND LCP bit-exact (stricter than the 1e-5 relative the north star asks for)."""
import numpy as np
import pytest

import oracle
from model_matching_b200 import Context, synth

pytestmark = pytest.mark.gpu


def _check(ctx, sc, mpos, mnrm, T):
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    cs, cm = ctx.centroids()
    ocs, ocm = est.centroids()
    assert np.array_equal(cs, ocs) and np.array_equal(cm, ocm)
    s, m = ctx.centred()
    os_, om = est.centred()
    assert np.array_equal(s, os_) and np.array_equal(m, om)
    lcp, inl = ctx.score_lcp(T)
    olcp, oinl = est.score(T, threads=8)
    assert np.array_equal(inl, oinl)
    assert np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32))
    return lcp, inl


def test_score_small_scene(gpu_ctx, small_scene):
    sc, mpos, mnrm = small_scene
    T, near = synth.make_hypotheses(4000, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=11, near_fraction=0.05)
    lcp, inl = _check(gpu_ctx, sc, mpos, mnrm, T)
    assert inl[near].min() > 100 and inl.max() <= len(mpos)


@pytest.mark.parametrize("M", [1, 31, 33, 500])
def test_score_ragged_model_sizes(gpu_ctx, small_scene, M):
    sc, _, _ = small_scene
    mpos, mnrm = synth.make_model(M)
    T, _ = synth.make_hypotheses(600, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=5, near_fraction=0.2)
    _check(gpu_ctx, sc, mpos, mnrm, T)


def test_score_exact_ties_use_kdtree_rule(gpu_ctx):
    """Scene points on an exact lattice, queries exactly between them: many exact d^2 ties whose
    winner decides the normal test.  The GPU must pick the kd-tree's winner."""
    g = np.arange(-8, 9, dtype=np.float32) * np.float32(0.0078125)   # 2^-7 spacing, exact
    X, Y = np.meshgrid(g, g, indexing="ij")
    pos = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size, np.float32)], -1)
    rng = np.random.default_rng(3)
    nrm = rng.normal(size=pos.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    cls = rng.integers(1000, 10001, size=len(pos)).astype(np.float32) / np.float32(10000)
    # symmetric scene => centroid exactly 0; model points at lattice midpoints (two or four
    # scene points at exactly the same distance)
    assert np.all(pos.sum(0) == 0)
    h = np.float32(0.00390625)
    mp_ = np.array([[h, 0, 0], [-h, 0, 0], [h, h, 0], [0, h, 0], [3 * h, h, 0], [h, -3 * h, 0]], np.float32)
    mp_ = np.concatenate([mp_, -mp_])   # centroid exactly 0
    mn = np.tile(np.array([[0, 0, 1]], np.float32), (len(mp_), 1))
    ctx = gpu_ctx
    ctx.upload_model(mp_, mn)
    ctx.upload_scene(pos, nrm, cls)
    # pure lattice translations keep the ties exact
    T = []
    for i in range(-4, 5):
        for j in range(-4, 5):
            M4 = np.eye(4, dtype=np.float32)
            M4[0, 3] = i * 2 * h
            M4[1, 3] = j * 2 * h
            T.append(M4.T.reshape(16))
    T = np.array(T, np.float32)
    est = oracle.Estimator(pos, nrm, cls, mp_, mn, distance_threshold=0.008)
    ctx2 = Context(0, distance_threshold=0.008)
    ctx2.upload_model(mp_, mn)
    ctx2.upload_scene(pos, nrm, cls)
    before = ctx2.counters()[1]
    lcp, inl = ctx2.score_lcp(T)
    olcp, oinl = est.score(T)
    assert ctx2.counters()[1] - before > 100, "the test must actually exercise the tie path"
    assert np.array_equal(inl, oinl)
    assert np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32))
    ctx2.close()


def test_score_empty_and_out_of_grid(gpu_ctx, small_scene):
    sc, mpos, mnrm = small_scene
    gpu_ctx.upload_model(mpos, mnrm)
    gpu_ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    lcp, inl = gpu_ctx.score_lcp(np.zeros((0, 16), np.float32))
    assert lcp.size == 0
    far = np.eye(4, dtype=np.float32)
    far[:3, 3] = [1e6, -1e6, 3e5]
    nanT = np.full(16, np.nan, np.float32)
    T = np.stack([far.T.reshape(16), nanT, np.zeros(16, np.float32)])
    lcp, inl = gpu_ctx.score_lcp(T)
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    olcp, oinl = est.score(T)
    assert np.array_equal(inl, oinl) and np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32))


def test_reduce_best_matches_reference_rule(gpu_ctx):
    rng = np.random.default_rng(0)
    lcp = rng.uniform(0, 1, 100000).astype(np.float32)
    lcp[rng.integers(0, lcp.size, 50)] = lcp.max()          # ties on the maximum: first wins
    bi, bl, ti, tl = gpu_ctx.reduce_best(lcp, K=32)
    assert (bi, bl) == oracle.best(lcp)
    order = np.lexsort((np.arange(lcp.size), -lcp.astype(np.float64)))[:32]
    assert np.array_equal(ti, order) and np.array_equal(tl, lcp[order])
    z = np.zeros(1000, np.float32)
    bi, bl, ti, tl = gpu_ctx.reduce_best(z, K=4)
    assert bi == -1 and bl == 0 and np.all(ti == -1)
    few = np.array([0, 0.5, 0, 0.25], np.float32)
    bi, bl, ti, tl = gpu_ctx.reduce_best(few, K=4)
    assert bi == 1 and list(ti) == [1, 3, -1, -1]


def test_backproject_bit_exact(gpu_ctx):
    rng = np.random.default_rng(1234)
    depth = rng.integers(0, 20000, size=(480, 640)).astype(np.uint16)
    depth[rng.uniform(size=depth.shape) < 0.1] = 0
    bgr = rng.integers(0, 256, size=(480, 640, 3)).astype(np.uint8)
    fx, cx, fy, cy = 572.4114, 325.2611, 573.57043, 242.04899
    xyz, rgb = gpu_ctx.backproject(depth, bgr, fx, cx, fy, cy, 1 / 1000.0)
    oxyz, orgb = oracle.backproject(depth, bgr, fx, cx, fy, cy, 1 / 1000.0)
    assert np.array_equal(xyz.view(np.uint32), oxyz.view(np.uint32))
    assert np.array_equal(rgb, orgb)
    xyz2, none = gpu_ctx.backproject(depth[:7, :5].copy(), None, fx, cx, fy, cy, 1 / 8000.0)
    assert none is None
    oxyz2, _ = oracle.backproject(depth[:7, :5].copy(), None, fx, cx, fy, cy, 1 / 8000.0)
    assert np.array_equal(xyz2.view(np.uint32), oxyz2.view(np.uint32))


def test_select_above_is_the_clustering_prefilter(gpu_ctx):
    rng = np.random.default_rng(4)
    lcp = rng.uniform(0, 1, 50000).astype(np.float32)
    lcp[rng.integers(0, lcp.size, 100)] = np.float32(0.75)    # equal scores: index ascending
    best = float(lcp.max())
    thr = np.float32(0.7) * np.float32(best)
    idx, val = gpu_ctx.select_above(lcp, float(thr))
    keep = np.flatnonzero(lcp > thr)
    order = keep[np.lexsort((keep, -lcp[keep].astype(np.float64)))]
    assert np.array_equal(idx, order) and np.array_equal(val, lcp[order])
    idx0, _ = gpu_ctx.select_above(np.zeros(10, np.float32), 0.0)
    assert idx0.size == 0


def test_multi_object_sweep_shares_one_scene_index(gpu_ctx, small_scene):
    """BASELINE.json configs[4] at test size: eight models (|M| in 384..1024) scored against the same
    scene; the scene index is built once, only the model tables are swapped."""
    sc, _, _ = small_scene
    gpu_ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    cells_before = gpu_ctx.counters()[2]
    for k, M in enumerate([384, 448, 512, 600, 640, 768, 900, 1024]):
        mpos, mnrm = synth.make_model(M, radius=0.05 + 0.005 * k)
        gpu_ctx.upload_model(mpos, mnrm)
        T, _ = synth.make_hypotheses(1500, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=1000 + k, near_fraction=0.05)
        est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
        lcp, inl = gpu_ctx.score_lcp(T)
        olcp, oinl = est.score(T, threads=8)
        assert np.array_equal(inl, oinl), M
        assert np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32)), M
    assert gpu_ctx.counters()[2] == cells_before


def test_host_call_pinned_and_pageable_buffers_agree(gpu_ctx, small_scene):
    """stocs_b200_score_lcp reads page-locked transforms in place (zero-copy, one launch) and
    stages pageable ones through HBM in chunks on two streams; both must give the same bits."""
    import torch
    sc, mpos, mnrm = small_scene
    H = 150_000                                   # above the 2^16 threshold of the chunked path
    T, _ = synth.make_hypotheses(H, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=5, near_fraction=0.03)
    gpu_ctx.upload_model(mpos, mnrm)
    gpu_ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    lcp_pg, inl_pg = gpu_ctx.score_lcp(T)        # numpy = pageable
    hT = torch.from_numpy(T).pin_memory()
    hl = torch.empty(H, dtype=torch.float32).pin_memory()
    hi = torch.empty(H, dtype=torch.int32).pin_memory()
    gpu_ctx.score_lcp_ptr(hT.data_ptr(), H, hl.data_ptr(), hi.data_ptr())
    assert np.array_equal(hl.numpy().view(np.uint32), lcp_pg.view(np.uint32))
    assert np.array_equal(hi.numpy(), inl_pg)
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    olcp, oinl = est.score(T[:20000], threads=8)
    assert np.array_equal(inl_pg[:20000], oinl) and np.array_equal(lcp_pg[:20000].view(np.uint32), olcp.view(np.uint32))
    # the reduction of the resident array (score_lcp computes it behind its result copies) is the
    # one an explicit array gives, for both paths and every K
    want = gpu_ctx.reduce_best(lcp_pg, K=32)
    assert (want[0], want[1]) == oracle.best(lcp_pg)
    gpu_ctx.score_lcp_ptr(hT.data_ptr(), H, hl.data_ptr(), hi.data_ptr())
    for K in (1, 7, 32):
        got = gpu_ctx.reduce_best(None, K=K)
        assert got[0] == want[0] and got[1] == want[1]
        assert np.array_equal(got[2], want[2][:K]) and np.array_equal(got[3], want[3][:K])
    gpu_ctx.score_lcp(T)
    got = gpu_ctx.reduce_best(None, K=32)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
    # a smaller second call must not be answered from the first call's cache
    lcp_s, _ = gpu_ctx.score_lcp(T[:70000])
    got = gpu_ctx.reduce_best(None, K=4)
    assert (got[0], got[1]) == oracle.best(lcp_s)


@pytest.mark.parametrize("scale,coarse_bits,max_cells", [("0.25", None, None), ("0.5", "2048", None), ("0.75", None, None),
                                                         ("1.0", "1024", None), ("2.0", None, None), ("3.7", None, None),
                                                         ("0.5", None, "1000000")])
def test_index_geometry_knobs_do_not_change_results(small_scene, scale, coarse_bits, max_cells, monkeypatch):
    """Cell edge (in eps) and coarse-map capacity are tuning knobs of the scene index: apron width,
    padding to whole coarse blocks and the always-empty border block must keep every setting
    bit-identical to the oracle (hypotheses reach up to eps outside the scene's bounding box)."""
    sc, mpos, mnrm = small_scene
    monkeypatch.setenv("STOCS_CELL_SCALE", scale)
    if coarse_bits:
        monkeypatch.setenv("STOCS_COARSE_BITS", coarse_bits)
    if max_cells:  # the cell edge is enlarged until the grid fits (the fallback for very large scenes)
        monkeypatch.setenv("STOCS_MAX_CELLS", max_cells)
    T, _ = synth.make_hypotheses(3000, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=21, near_fraction=0.05)
    # poses that put model points just outside every face of the scene's bounding box
    lo, hi = sc["pos"].min(0), sc["pos"].max(0)
    ts = []
    for axis in range(3):
        for side, edge in ((-1, lo), (1, hi)):
            for off in (0.0, 0.0049, 0.0051, 0.02):
                c = (lo + hi) / 2
                c[axis] = edge[axis] + side * off
                ts.append(c - mpos.mean(0))
    ts = np.array(ts, np.float32)
    extra = synth.to_colmajor16(np.tile(np.eye(3, dtype=np.float32), (len(ts), 1, 1)), ts)
    T = np.concatenate([T.reshape(-1, 16), extra.reshape(-1, 16)]).reshape(-1, 16)
    ctx = Context(0)
    try:
        _check(ctx, sc, mpos, mnrm, T)
    finally:
        ctx.close()


def test_centroid_host_and_device_paths_agree(small_scene):
    """the order-dependent fp32 centroid (src/stocs.cpp:945-956) is summed on the host while the
    points are in flight; the single-CTA device kernel (STOCS_DEVICE_CENTROID=1) must give the same bits"""
    import os
    from model_matching_b200 import Context
    sc, mpos, mnrm = small_scene
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    want_cs, want_cm = est.centroids()
    want_s, want_m = est.centred()
    for dev_path in (False, True):
        if dev_path:
            os.environ["STOCS_DEVICE_CENTROID"] = "1"
        try:
            ctx = Context(0)
            ctx.upload_model(mpos, mnrm)
            ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
        finally:
            os.environ.pop("STOCS_DEVICE_CENTROID", None)
        cs, cm = ctx.centroids()
        s, m = ctx.centred()
        assert np.array_equal(cs.view(np.uint32), want_cs.view(np.uint32)) and np.array_equal(cm.view(np.uint32), want_cm.view(np.uint32))
        assert np.array_equal(s.view(np.uint32), want_s.view(np.uint32)) and np.array_equal(m.view(np.uint32), want_m.view(np.uint32))
        ctx.close()


def test_heavy_first_schedule_changes_no_result(small_scene, monkeypatch):
    """launches of 32 768..300 000 hypotheses are claimed in a heavy-first order (probe_order_kernel);
    the order only decides which warp scores a hypothesis when"""
    sc, mpos, mnrm = small_scene
    T, _ = synth.make_hypotheses(40000, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=21, near_fraction=0.02)
    ctx = Context(0)
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    lcp1, inl1 = ctx.score_lcp(T)
    monkeypatch.setenv("STOCS_NO_LPT", "1")
    lcp0, inl0 = ctx.score_lcp(T)
    monkeypatch.delenv("STOCS_NO_LPT")
    monkeypatch.setenv("STOCS_LPT_MAX", "10000000")       # also on the largest size
    lcp2, inl2 = ctx.score_lcp(np.concatenate([T] * 9))
    assert np.array_equal(inl1, inl0) and np.array_equal(lcp1.view(np.uint32), lcp0.view(np.uint32))
    assert np.array_equal(inl2, np.tile(inl0, 9)) and np.array_equal(lcp2.view(np.uint32), np.tile(lcp0, 9).view(np.uint32))
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    olcp, oinl = est.score(T[:3000], threads=4)
    assert np.array_equal(inl1[:3000], oinl) and np.array_equal(lcp1[:3000].view(np.uint32), olcp.view(np.uint32))
    ctx.close()


def test_long_thin_scene_cell_location_margin():
    """a 60 m strip at 2.5 mm cells is 24 000 cells long: the FMA-evaluated cell map is off by more
    than cell/256 at the far end, so the candidate-list margin scales with the grid extent"""
    rng = np.random.default_rng(5)
    n = 60000
    pos = np.stack([rng.uniform(0, 60.0, n), rng.uniform(0, 0.04, n), rng.uniform(0, 0.04, n)], -1).astype(np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (n, 1))
    cls = rng.uniform(0.1, 1, n).astype(np.float32)
    mpos = rng.uniform(-0.02, 0.02, (128, 3)).astype(np.float32)
    mnrm = np.tile(np.array([0, 0, 1], np.float32), (128, 1))
    ctx = Context(0)
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(pos, nrm, cls)
    cs, cm = ctx.centroids()
    H = 4000
    R = synth.axis_angle(rng.normal(size=(H, 3)), rng.uniform(0, 0.3, H))
    t = np.stack([rng.uniform(0, 60.0, H), rng.uniform(0, 0.04, H), rng.uniform(0, 0.04, H)], -1)
    t[:H // 2, 0] = rng.uniform(59.0, 60.0, H // 2)          # half of them at the far end of the grid
    T = synth.to_colmajor16(R, t - cs.astype(np.float64))
    lcp, inl = ctx.score_lcp(T)
    est = oracle.Estimator(pos, nrm, cls, mpos, mnrm)
    olcp, oinl = est.score(T, threads=8)
    assert oinl.sum() > 1000
    assert np.array_equal(inl, oinl) and np.array_equal(lcp.view(np.uint32), olcp.view(np.uint32))
    ctx.close()
