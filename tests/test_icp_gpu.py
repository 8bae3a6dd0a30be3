"""GPU point-to-plane ICP (csrc/icp.cu) against the oracle: bit-identical transform, moved source
and pair counts (the binary64 sums are taken in the same fixed order on both sides)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from model_matching_b200 import synth
from test_icp import _cloud

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from model_matching_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32), np.ascontiguousarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("n_src,n_tgt", [(1, 40), (3, 40), (255, 300), (256, 300), (257, 1024), (5000, 1025), (1237, 3100)])
def test_icp_bit_exact_against_oracle(ctx, n_src, n_tgt):
    tgt, tn = _cloud(n_tgt, 10 + n_tgt)
    rng = np.random.default_rng(n_src)
    R = synth.axis_angle(rng.normal(size=3), 0.05)
    pick = rng.integers(0, n_tgt, n_src)
    src = (tgt[pick] @ R.T + rng.normal(scale=0.002, size=(n_src, 3))).astype(np.float32)
    src[::7] += np.float32(0.5)                 # a share of the source finds no partner
    want = oracle.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    got = ctx.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    assert _same(got[0], want[0])
    assert _same(got[1], want[1])
    assert np.array_equal(got[2], want[2]) and got[3] == want[3] and got[4] == want[4]


def test_icp_ties_and_lattice_targets(ctx):
    # lattice target: many source points are exactly equidistant from several targets
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * np.float32(0.0078125)
    tn = np.tile(np.array([0, 0.6, 0.8], np.float32), (len(g), 1))
    src = (g[:-1] + g[1:]) * np.float32(0.5)    # midpoints: exact ties in binary32
    want = oracle.icp_point_to_plane(src, g, tn, 3, 0.035)
    got = ctx.icp_point_to_plane(src, g, tn, 3, 0.035)
    assert _same(got[0], want[0]) and _same(got[1], want[1]) and np.array_equal(got[2], want[2])


def test_icp_not_converged_and_iteration_counts(ctx):
    tgt, tn = _cloud(300, 3)
    src = tgt[:50] + np.float32(1.0)
    T, moved, pairs, done, conv = ctx.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    assert not conv and done == 0 and pairs[0] == 0
    assert np.array_equal(T, np.eye(4, dtype=np.float32)) and np.array_equal(moved, src)
    for iters in (1, 2, 16):
        want = oracle.icp_point_to_plane(tgt[::3] + np.float32(0.001), tgt, tn, iters, 0.035)
        got = ctx.icp_point_to_plane(tgt[::3] + np.float32(0.001), tgt, tn, iters, 0.035)
        assert _same(got[0], want[0]) and got[3] == want[3] == iters
    with pytest.raises(Exception):
        ctx.icp_point_to_plane(src, tgt, tn, 17, 0.035)
    with pytest.raises(Exception):
        ctx.icp_point_to_plane(src, tgt, tn, 5, 0.0)


def test_host_point_to_plane_icp_wrapper():
    """clustering::point_to_plane_icp (host/pose_clustering.cpp) through the test CLI"""
    exe = os.path.join(ROOT, "model_matching_b200", "host", "test_clustering")
    tgt, tn = _cloud(600, 5)
    R = synth.axis_angle(np.array([0.2, 0.9, -0.1]), 0.04)
    src = (tgt[::3] @ R.T + np.array([0.001, 0.002, -0.001])).astype(np.float32)
    text = "%d %d\n" % (len(src), len(tgt))
    text += "\n".join("%.9g %.9g %.9g" % tuple(p) for p in src) + "\n"
    text += "\n".join("%.9g %.9g %.9g %.9g %.9g %.9g" % (*p, *n) for p, n in zip(tgt, tn)) + "\n"
    out = subprocess.run([exe, "icp"], input=text, capture_output=True, text=True, timeout=120, check=True).stdout.split("\n")
    T = np.array(out[0].split(), np.float32).reshape(4, 4).T
    moved = np.array([l.split() for l in out[1:1 + len(src)]], np.float32)
    want = oracle.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    assert _same(T, want[0]) and _same(moved, want[1])


def test_host_icp_wrapper_resets_offset_when_not_converged():
    """src/pose_clustering.cpp:136-139: the caller's matrix becomes identity when PCL's ICP has not
    converged -- checked from a NON-identity starting matrix (a stale value must not survive)."""
    exe = os.path.join(ROOT, "model_matching_b200", "host", "test_clustering")
    tgt, tn = _cloud(300, 3)
    src = tgt[:50] + np.float32(1.0)            # nothing within the correspondence distance
    text = "%d %d\n" % (len(src), len(tgt))
    text += "\n".join("%.9g %.9g %.9g" % tuple(p) for p in src) + "\n"
    text += "\n".join("%.9g %.9g %.9g %.9g %.9g %.9g" % (*p, *n) for p, n in zip(tgt, tn)) + "\n"
    out = subprocess.run([exe, "icp", "stale"], input=text, capture_output=True, text=True, timeout=120, check=True).stdout.split("\n")
    T = np.array(out[0].split(), np.float32).reshape(4, 4).T
    assert np.array_equal(T, np.eye(4, dtype=np.float32))
    # and a converging run from the same stale matrix returns the ICP transform, not a product with it
    R = synth.axis_angle(np.array([0.2, 0.9, -0.1]), 0.04)
    src2 = (tgt[::3] @ R.T).astype(np.float32)
    text = "%d %d\n" % (len(src2), len(tgt))
    text += "\n".join("%.9g %.9g %.9g" % tuple(p) for p in src2) + "\n"
    text += "\n".join("%.9g %.9g %.9g %.9g %.9g %.9g" % (*p, *n) for p, n in zip(tgt, tn)) + "\n"
    out = subprocess.run([exe, "icp", "stale"], input=text, capture_output=True, text=True, timeout=120, check=True).stdout.split("\n")
    T2 = np.array(out[0].split(), np.float32).reshape(4, 4).T
    assert _same(T2, oracle.icp_point_to_plane(src2, tgt, tn, 5, 0.035)[0])
