"""GPU parity of the compact PPF table against the reference's fully expanded std::map
(include/rgbd.hpp:23, src/rgbd.cpp:123-154, src/stocs.cpp:62-78) as restated by the oracle."""
import numpy as np
import pytest

import oracle
from model_matching_b200 import synth

pytestmark = pytest.mark.gpu


def _model(n, seed=0):
    rng = np.random.default_rng(seed)
    pos, nrm = synth.make_model(n, radius=0.06)
    # perturb normals a little so that angle bins are not degenerate
    nrm = nrm + rng.normal(0, 0.2, nrm.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pos = pos + np.array([0.3, -0.2, 0.9], np.float32)   # un-centred, like a file model
    return pos.astype(np.float32), nrm.astype(np.float32)


def test_ppf_table_matches_expanded_map(gpu_ctx):
    pos, nrm = _model(150)
    omap = oracle.PPFMap(pos, nrm)
    gpu_ctx.upload_model(pos, nrm)
    npairs, nbins = gpu_ctx.ppf_num_pairs()
    assert npairs == 150 * 149
    keys = omap.keys()
    rng = np.random.default_rng(1)
    pick = rng.choice(len(keys), size=400, replace=False)
    for k in keys[pick]:
        want = omap.lookup(k)
        got = gpu_ctx.ppf_lookup(k)
        assert got is not None and np.array_equal(got, want), k
    # absent keys: perturb existing ones and keep those the map does not hold
    present = {tuple(k) for k in keys}
    n_absent = 0
    for k in keys[pick]:
        for d in ([5, 0, 0, 0], [0, 20, 0, 0], [0, 0, -20, 0], [-40, 0, 0, 20], [1, 0, 0, 0]):
            q = k + np.array(d, np.int32)
            if tuple(q) in present:
                continue
            assert gpu_ctx.ppf_lookup(q) is None, q
            n_absent += 1
    assert n_absent > 100
    for q in ([5, 0, 0, 0], [0, 0, 0, 0], [-5, 10, 10, 10], [10, -5, 0, 0], [100000, 0, 0, 0]):
        assert gpu_ctx.ppf_lookup(np.array(q, np.int32)) is None


def test_preloaded_table_round_trip_and_rejection(gpu_ctx):
    """stocs_b200_upload_ppf_table (reference: ppf_map = ppf_map_preloaded, src/stocs.cpp:94): the
    exported table uploaded again gives the same lookups; a thinned table is what gets used; a table
    of another discretisation, another model size or with entries off the lattice is refused"""
    from model_matching_b200 import StocsError
    pos, nrm = _model(120)
    gpu_ctx.upload_model(pos, nrm)
    keys, pairs = gpu_ctx.ppf_export()
    n_keys = gpu_ctx.ppf_num_expanded_keys()
    probe = keys[:: max(1, len(keys) // 60)]
    before = [gpu_ctx.ppf_lookup(k) for k in probe]
    perm = np.random.default_rng(3).permutation(len(keys))          # any entry order is accepted
    gpu_ctx.upload_ppf_table(keys[perm], pairs[perm], 5, 5, 120)
    assert gpu_ctx.ppf_num_expanded_keys() == n_keys and gpu_ctx.ppf_num_pairs()[0] == len(pairs)
    for k, want in zip(probe, before):
        assert np.array_equal(gpu_ctx.ppf_lookup(k), want)
    keep = pairs[:, 0] % 3 == 0
    gpu_ctx.upload_ppf_table(keys[keep], pairs[keep], 5, 5, 120)
    assert gpu_ctx.ppf_num_pairs()[0] == int(keep.sum())
    for k, want in zip(probe, before):
        got = gpu_ctx.ppf_lookup(k)
        w = want[want[:, 0] % 3 == 0]
        assert (got is None and len(w) == 0) or np.array_equal(got, w)
    for bad_args in ((keys, pairs, 10, 5, 120), (keys, pairs, 5, 5, 121)):
        with pytest.raises(StocsError):
            gpu_ctx.upload_ppf_table(*bad_args)
        gpu_ctx.upload_model(pos, nrm)          # a refused table leaves the context without a model
    off = keys.copy(); off[0, 1] += 1           # not a multiple of the rotation bin
    with pytest.raises(StocsError):
        gpu_ctx.upload_ppf_table(off, pairs, 5, 5, 120)
    gpu_ctx.upload_model(pos, nrm)
    big = pairs.copy(); big[0, 1] = 120         # id out of range
    with pytest.raises(StocsError):
        gpu_ctx.upload_ppf_table(keys, big, 5, 5, 120)
