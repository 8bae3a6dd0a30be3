"""Pins of the oracle's restated data structures against brute force / hand-derived answers:
kd-tree NN (kdtree.h:394-459), the expanded PPF map (rgbd.cpp:123-154), rigid fit
(stocs.cpp:270-361), base ordering (stocs.cpp:224-268), congruent sets (stocs.cpp:753-869)."""
import numpy as np

import oracle
from model_matching_b200 import synth
from scenes import object_scene


def test_kdtree_matches_brute_force(small_scene):
    sc, mpos, mnrm = small_scene
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    s, _ = est.centred()
    rng = np.random.default_rng(5)
    q = rng.uniform(s.min(0), s.max(0), size=(4000, 3)).astype(np.float32)
    q[:2000] = s[rng.integers(0, len(s), 2000)] + rng.normal(0, 0.003, (2000, 3)).astype(np.float32)
    eps2 = np.float32(0.005) ** 2
    ids = est.kd_query(q, eps2)
    # same fp32 distance formula as the tree: dx*dx + (dy*dy + dz*dz)
    d = q[:, None, :] - s[None, :, :]
    d2 = d[..., 0] * d[..., 0] + (d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    best = d2.argmin(1)
    found = d2.min(1) <= eps2
    assert found.sum() > 500
    assert np.array_equal(ids >= 0, found)
    assert np.array_equal(ids[found], best[found])


def test_kdtree_inclusive_radius_and_ties():
    pos = np.array([[0, 0, 0], [0.01, 0, 0], [-0.01, 0, 0], [0.5, 0.5, 0.5]], np.float32)
    pos = pos - pos.mean(0, dtype=np.float64).astype(np.float32)  # keep it nearly centred
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    est = oracle.Estimator(pos, nrm, np.ones(4, np.float32), pos[:2], nrm[:2])
    s, _ = est.centred()
    q = s[0:1] + np.array([[0.005, 0, 0]], np.float32)
    d2 = float(((q - s[0]) ** 2).sum())
    assert est.kd_query(q, np.float32(d2))[0] in (0, 1)          # on the radius: accepted (<=)
    assert est.kd_query(q, np.float32(d2 * 0.49))[0] == -1       # inside the gap: nothing


def test_ppf_map_expansion_rule():
    """Every ordered pair is inserted under keys {f1-5, f1} x {f-10, f-5, f, f+5}^3 minus those with
    p1 <= 5 or a negative angle (rgbd.cpp:130-137); lists are in (id1, id2) order."""
    rng = np.random.default_rng(2)
    pos = rng.uniform(-0.05, 0.05, (40, 3)).astype(np.float32)
    nrm = rng.normal(size=(40, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    m = oracle.PPFMap(pos, nrm)
    want = {}
    for i in range(40):
        for j in range(40):
            if i == j:
                continue
            f = oracle.ppf_compute(pos[i:i + 1], nrm[i:i + 1], pos[j:j + 1], nrm[j:j + 1])[0]
            for a in (f[0] - 5, f[0]):
                for b in range(f[1] - 10, f[1] + 10, 5):
                    for c in range(f[2] - 10, f[2] + 10, 5):
                        for d in range(f[3] - 10, f[3] + 10, 5):
                            if a <= 5 or b < 0 or c < 0 or d < 0:
                                continue
                            want.setdefault((a, b, c, d), []).append((i, j))
    assert m.num_keys == len(want)
    assert m.num_entries == sum(len(v) for v in want.values())
    keys = m.keys()
    assert [tuple(k) for k in keys] == sorted(want)               # std::map order
    for k in keys[:: max(1, len(keys) // 200)]:
        assert m.lookup(k).tolist() == [list(p) for p in want[tuple(k)]]
    assert m.lookup(np.array([5, 0, 0, 0], np.int32)) is None


def test_fit_recovers_a_planted_rigid_motion():
    rng = np.random.default_rng(3)
    mpos = rng.uniform(-0.1, 0.1, (50, 3)).astype(np.float32)
    mnrm = np.tile(np.array([[0, 0, 1]], np.float32), (50, 1))
    R = synth.axis_angle(np.array([0.2, 1.0, -0.4]), 0.9)
    t = np.array([0.3, -0.2, 0.8])
    spos = (mpos.astype(np.float64) @ R.T + t).astype(np.float32)
    est = oracle.Estimator(spos, mnrm, np.ones(50, np.float32), mpos, mnrm)
    ok, Tc, Tw = est.fit(np.array([3, 17, 29, 41], np.int32), np.array([3, 17, 29, 41], np.int32))
    assert ok
    Tw = Tw.reshape(4, 4).T
    assert np.allclose(Tw[:3, :3], R, atol=2e-5) and np.allclose(Tw[:3, 3], t, atol=2e-5)
    assert np.allclose(Tw[3], [0, 0, 0, 1])
    # centred transform maps centred model onto centred scene
    s, m = est.centred()
    Tc = Tc.reshape(4, 4).T
    assert np.allclose(m @ Tc[:3, :3].T + Tc[:3, 3], s, atol=2e-5)
    # degenerate input is rejected (deviation D1)
    assert not est.fit(np.array([3, 3, 29, 41], np.int32), np.array([3, 17, 29, 41], np.int32))[0]
    assert not est.fit(np.array([3, 17, 29, 41], np.int32), np.array([5, 5, 5, 7], np.int32))[0]


def test_try_sampled_base_orders_crossing_segments():
    # a planar quad whose diagonals cross at known ratios: p0-p1 and p2-p3 intersect at
    # 0.25 along the first and 0.5 along the second segment
    pos = np.array([[0, 0, 0], [0.4, 0, 0], [0.1, -0.1, 0], [0.1, 0.1, 0], [0.9, 0.9, 0.3]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (5, 1))
    est = oracle.Estimator(pos, nrm, np.ones(5, np.float32), pos[:3], nrm[:3])
    for perm in ([0, 1, 2, 3], [2, 0, 3, 1], [3, 2, 1, 0]):
        ok, ids, inv = est.try_sampled_base(np.array(perm, np.int32))
        assert ok and sorted(ids.tolist()) == [0, 1, 2, 3]
        seg1, seg2 = tuple(ids[:2]), tuple(ids[2:])
        assert {frozenset(seg1), frozenset(seg2)} == {frozenset((0, 1)), frozenset((2, 3))}
        want = {(0, 1): 0.25, (1, 0): 0.75, (2, 3): 0.5, (3, 2): 0.5}
        assert abs(inv[0] - want[seg1]) < 1e-5 and abs(inv[1] - want[seg2]) < 1e-5


def test_congruent_sets_contain_the_planted_correspondence():
    """A base sampled on the planted instance must list, among its congruent quads, model
    quadrilaterals whose fitted transform explains the model (high inlier count)."""
    sc, mpos, mnrm = object_scene()
    omap = oracle.PPFMap(mpos, mnrm)
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm, ppfmap=omap)
    best = 0
    n_valid = 0
    for b in range(24):
        ok, ids, inv, stage = est.sample_class_base(20181018, b)
        if not ok:
            continue
        n_valid += 1
        quads, nP, nQ = est.find_congruent(ids, inv[0], inv[1])
        # set order: strictly increasing (rank in P, rank in Q) == lexicographic in the quad when
        # P and Q lists are (id1,id2)-sorted
        if len(quads) > 1:
            pq = [tuple(q) for q in quads]
            assert pq == sorted(pq)
        Ts = [est.fit(ids, q)[1] for q in quads[:60]]
        if Ts:
            lcp, inl = est.score(np.array(Ts))
            best = max(best, int(inl.max()))
    assert n_valid >= 5
    assert best > 0.6 * len(mpos)


def test_sampling_is_reproducible_and_seed_dependent():
    sc, mpos, mnrm = object_scene()
    omap = oracle.PPFMap(mpos, mnrm)
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm, ppfmap=omap)
    a = [est.sample_class_base(1, b)[1].tolist() for b in range(8)]
    b_ = [est.sample_class_base(1, b)[1].tolist() for b in range(8)]
    c = [est.sample_class_base(2, b)[1].tolist() for b in range(8)]
    assert a == b_ and a != c
    # surviving probabilities after the last pass are either 0 or the class prior
    est.sample_class_base(1, 0)
    cur = est.current_prob()
    assert np.all((cur == 0) | (cur == sc["cls"]))


def test_score_matches_an_independent_numpy_restatement(small_scene):
    """compute_alignment_score_for_rigid_transform (src/stocs.cpp:1006-1041) written a second time,
    in vectorised numpy float32 with brute-force NN instead of the kd-tree and numpy's arccos instead
    of the shared math header: inlier counts must agree exactly, LCP to 1e-6 relative."""
    sc, mpos, mnrm = small_scene
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    T, near = synth.make_hypotheses(60, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=21, near_fraction=0.5)
    lcp, inl = est.score(T)
    s, m = est.centred()
    f = np.float32
    eps2 = f(0.005) * f(0.005)
    for h in range(len(T)):
        A = T[h].reshape(4, 4).T.astype(np.float32)
        q = np.stack([((A[r, 0] * m[:, 0] + A[r, 1] * m[:, 1]) + A[r, 2] * m[:, 2]) + A[r, 3] for r in range(3)], -1)
        nq = np.stack([A[r, 0] * mnrm[:, 0] + (A[r, 1] * mnrm[:, 1] + A[r, 2] * mnrm[:, 2]) for r in range(3)], -1)
        count, acc = 0, f(0)
        # brute force only near the scene (all model points whose box is close); chunked for memory
        for i in range(len(m)):
            d = q[i] - s
            d2 = d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
            j = int(np.argmin(d2))
            if d2[j] <= eps2:
                n = sc["nrm"][j]
                dot = n[0] * nq[i, 0] + (n[1] * nq[i, 1] + n[2] * nq[i, 2])
                if abs(dot) <= 1:
                    ang = f(np.float64(f(f(np.arccos(np.float64(dot))) * f(180))) / np.pi)
                    if ang < 30:
                        count += 1
                        acc = f(acc + sc["cls"][j])
        assert count == inl[h], h
        assert abs(float(acc / f(len(m))) - float(lcp[h])) <= 1e-6 * max(1e-6, float(lcp[h])), h
    assert inl[near].min() > 50
