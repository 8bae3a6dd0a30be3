"""The reference kd-tree (kdtree.h:461-538) decides which of two scene points at EXACTLY the same distance
a query returns, through the order its partition leaves the points in.  The library builds it on the host
with a branch-free rewrite of the reference's partition loop; here that build is compared, on the CPU
alone, with the oracle's literal restatement: same leaf order of the original indices, same node count."""
import numpy as np
import pytest

import oracle
from model_matching_b200 import _capi
from scenes import object_scene


def _compare(pos, nrm, cls, mpos, mnrm):
    est = oracle.Estimator(pos, nrm, cls, mpos, mnrm)
    centred, _ = est.centred()
    want, want_nodes = est.kd_order()
    got, got_nodes = _capi.host_kdtree_order(centred)
    assert got_nodes == want_nodes
    assert np.array_equal(got, want)
    return want_nodes


def test_host_kdtree_matches_oracle_on_scenes(small_scene):
    sc, mpos, mnrm = small_scene
    assert _compare(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm) > 500
    sc2, mpos2, mnrm2 = object_scene()
    assert _compare(sc2["pos"], sc2["nrm"], sc2["cls"], mpos2, mnrm2) > 20


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_host_kdtree_matches_oracle_with_repeated_coordinates(seed):
    """coordinates on a coarse lattice: many values equal to each other and to the split value, whole
    nodes of identical points (depth cap), every branch of the partition's boundary handling"""
    rng = np.random.default_rng(seed)
    n = 5000
    pos = (rng.integers(0, 12, (n, 3)) * 0.0625).astype(np.float32)
    pos[: n // 5] = pos[0]                                   # 1000 identical points
    nrm = np.tile(np.array([0, 0, 1], np.float32), (n, 1))
    cls = np.ones(n, np.float32)
    mpos = rng.normal(0, 0.05, (64, 3)).astype(np.float32)
    mnrm = np.tile(np.array([0, 0, 1], np.float32), (64, 1))
    _compare(pos, nrm, cls, mpos, mnrm)
