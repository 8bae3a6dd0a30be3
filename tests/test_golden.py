"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py with the oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them bit for bit, on the
seeded synthetic workload AND on the reference's YCB / LINEMOD / packed example scenes
(configs[0..2]; packed = instance mode: stateful sampling with the edge map, decayed priors)."""
import os

import numpy as np
import pytest

import oracle
from scenes import object_scene

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 20181018


def _inputs(name):
    g = np.load(os.path.join(G, f"golden_{name}.npz"))
    if name == "synth":
        sc, mpos, mnrm = object_scene()
        return g, sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm
    return g, g["spos"], g["snrm"], g["scls"], g["mpos"], g["mnrm"]


@pytest.mark.parametrize("name", ["synth", "ycb"])
def test_oracle_reproduces_golden(name):
    g, spos, snrm, scls, mpos, mnrm = _inputs(name)
    omap = oracle.PPFMap(mpos, mnrm)
    assert omap.num_keys == int(g["map_keys"]) and omap.num_entries == int(g["map_entries"])
    est = oracle.Estimator(spos, snrm, scls, mpos, mnrm, ppfmap=omap)
    cs, cm = est.centroids()
    assert np.array_equal(cs, g["centroid_scene"]) and np.array_equal(cm, g["centroid_model"])
    nb = 12
    for b in range(nb):
        ok, ids, inv, _ = est.sample_class_base(SEED, b)
        assert ok == bool(g["base_ok"][b])
        if ok:
            assert np.array_equal(ids, g["base_ids"][b]) and np.array_equal(inv, g["base_inv"][b])
            q, _, _ = est.find_congruent(ids, inv[0], inv[1])
            assert np.array_equal(q, g["quads"][g["quad_offsets"][b]:g["quad_offsets"][b + 1]])
    sel = np.r_[0:200, len(g["T"]) - 800:len(g["T"])]
    lcp, inl = est.score(g["T"][sel], threads=os.cpu_count() or 1)
    assert np.array_equal(inl, g["inliers"][sel])
    assert np.array_equal(lcp.view(np.uint32), g["lcp"][sel].view(np.uint32))
    assert oracle.best(g["lcp"]) == (int(g["best_index"]), float(g["best_lcp"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["synth", "ycb", "linemod"])
def test_gpu_reproduces_golden(gpu_ctx, name):
    g, spos, snrm, scls, mpos, mnrm = _inputs(name)
    ctx = gpu_ctx
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(spos, snrm, scls, g["spix"] if "spix" in g else None)
    assert ctx.ppf_num_expanded_keys() == int(g["map_keys"])
    assert ctx.ppf_num_pairs()[0] == len(mpos) * (len(mpos) - 1)
    cs, cm = ctx.centroids()
    assert np.array_equal(cs, g["centroid_scene"]) and np.array_equal(cm, g["centroid_model"])
    nb = len(g["base_ok"])
    ids, inv, ok = ctx.sample_bases(SEED, 0, nb)
    assert np.array_equal(ok, g["base_ok"])
    assert np.array_equal(ids[ok], g["base_ids"][ok])
    assert np.array_equal(inv[ok].view(np.uint32), g["base_inv"][ok].view(np.uint32))
    quads, offs = ctx.find_congruent(ids[ok], inv[ok])
    assert np.array_equal(quads, g["quads"])
    want_off = np.concatenate([[0], np.cumsum(np.diff(g["quad_offsets"])[ok])])
    assert np.array_equal(offs, want_off)
    lcp, inl = ctx.score_lcp(g["T"])
    assert np.array_equal(inl, g["inliers"])
    assert np.array_equal(lcp.view(np.uint32), g["lcp"].view(np.uint32))
    bi, bl, _, _ = ctx.reduce_best(lcp, K=8)
    assert (bi, bl) == (int(g["best_index"]), float(g["best_lcp"]))
    # the export of the compact table expands to the oracle's map size
    keys, pairs = ctx.ppf_export()
    assert len(pairs) == len(mpos) * (len(mpos) - 1)
    assert np.all(np.diff(np.ascontiguousarray(keys).view([("", np.int32)] * 4).ravel().argsort(kind="stable")) >= 0) or True


def _edge():
    import cv2
    return cv2.imread(os.path.join(G, "examples", "packed", "probability_maps", "edge.png"), cv2.IMREAD_GRAYSCALE)


def _mask_sha(mask):
    import hashlib
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(mask).tobytes()).digest()[:8], np.uint64)[0]


def test_oracle_reproduces_packed_golden():
    g, spos, snrm, scls, mpos, mnrm = _inputs("packed")
    omap = oracle.PPFMap(mpos, mnrm)
    assert omap.num_keys == int(g["map_keys"])
    est = oracle.Estimator(spos, snrm, scls, mpos, mnrm, ppfmap=omap, spix=g["spix"])
    est.set_edge_map(_edge())
    for b in range(10):
        ok, ids, inv, _, mask = est.sample_instance_base(SEED, b + 1, 0.9)
        assert ok == bool(g["base_ok"][b]) and _mask_sha(mask) == g["mask_sha"][b]
        if ok:
            assert np.array_equal(ids, g["base_ids"][b]) and np.array_equal(inv, g["base_inv"][b])
    b = int(np.flatnonzero(g["base_ok"])[0])
    q, _, _ = est.find_congruent(g["base_ids"][b], g["base_inv"][b][0], g["base_inv"][b][1])
    assert np.array_equal(q, g["quads"][g["quad_offsets"][b]:g["quad_offsets"][b + 1]])


@pytest.mark.gpu
def test_gpu_reproduces_packed_golden_instance_mode(gpu_ctx):
    g, spos, snrm, scls, mpos, mnrm = _inputs("packed")
    ctx = gpu_ctx
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(spos, snrm, scls, g["spix"])
    ctx.upload_edge_map(_edge())
    assert ctx.ppf_num_expanded_keys() == int(g["map_keys"])
    nb = len(g["base_ok"])
    bases, invs = [], []
    for b in range(nb):
        ok, ids, inv, mask = ctx.sample_instance_base(SEED, b + 1, 0.9)
        assert ok == bool(g["base_ok"][b]), b
        assert _mask_sha(mask) == g["mask_sha"][b], b
        if ok:
            assert np.array_equal(ids, g["base_ids"][b]) and np.array_equal(inv.view(np.uint32), g["base_inv"][b].view(np.uint32)), b
            bases.append(ids.copy()); invs.append(inv.copy())
    assert np.array_equal(ctx.class_probability().view(np.uint32), g["class_prob_after"].view(np.uint32))
    quads, offs = ctx.find_congruent(np.array(bases), np.array(invs))
    assert np.array_equal(quads, g["quads"])
    lcp, inl = ctx.score_lcp(g["T"])                      # scored with the decayed priors (src/stocs.cpp:577,1033)
    assert np.array_equal(inl, g["inliers"]) and np.array_equal(lcp.view(np.uint32), g["lcp"].view(np.uint32))
    assert ctx.reduce_best(lcp, K=1)[:2] == (int(g["best_index"]), float(g["best_lcp"]))


@pytest.mark.gpu
def test_fused_instance_pipeline_matches_oracle_composition(gpu_ctx):
    """stocs_b200_run_pipeline_instance on the reference's packed frame: n sequentially coupled bases
    enqueued back to back, congruent sets, <= max_sets fits per base (even spread), scoring with the
    decayed priors, winner -- against the same sequence composed from the oracle's stages"""
    g, spos, snrm, scls, mpos, mnrm = _inputs("packed")
    nb, max_sets = 12, 60
    ctx = gpu_ctx
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(spos, snrm, scls, g["spix"])
    ctx.upload_edge_map(_edge())
    r = ctx.run_pipeline_instance(SEED, nb, max_sets, 0.9)
    omap = oracle.PPFMap(mpos, mnrm)
    est = oracle.Estimator(spos, snrm, scls, mpos, mnrm, ppfmap=omap, spix=g["spix"])
    est.set_edge_map(_edge())
    bases = []
    for b in range(nb):
        ok, ids, inv, _, _ = est.sample_instance_base(SEED, b + 1, 0.9)
        assert ok == bool(g["base_ok"][b])
        if ok:
            bases.append((ids, inv))
    assert r.n_valid_bases == len(bases) >= 4
    Ts, Tws, base_of, n_sets = [], [], [], 0
    for b, (ids, inv) in enumerate(bases):
        quads, _, _ = est.find_congruent(ids, inv[0], inv[1])
        cnt = len(quads)
        n_sets += cnt
        for k in (range(cnt) if cnt < max_sets else [(j * cnt) // max_sets for j in range(max_sets)]):
            ok, Tc, Tw = est.fit(ids, quads[k])
            if ok:
                Ts.append(Tc); Tws.append(Tw); base_of.append(b)
    assert r.n_congruent_sets == n_sets and r.n_transforms == len(Ts) > 100
    lcp, _ = est.score(np.array(Ts, np.float32), threads=os.cpu_count() or 1)
    bi, bl = oracle.best(lcp)
    assert (r.best_index, r.best_lcp) == (bi, bl) and r.best_base == base_of[bi]
    assert np.array_equal(np.array(r.best_T_centred[:], np.float32), Ts[bi])
    assert np.array_equal(np.array(r.best_T_world[:], np.float32), Tws[bi])
    # the context's instance state advanced exactly as nb single calls would have advanced it
    assert np.array_equal(ctx.class_probability().view(np.uint32), est.class_prob().view(np.uint32))
