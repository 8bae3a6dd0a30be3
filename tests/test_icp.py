"""Point-to-plane ICP (reference src/pose_clustering.cpp:123-141): the oracle against an independent
numpy restatement of PCL's linearised point-to-plane step, and the properties the loop must have."""
import numpy as np
import pytest

import oracle
from model_matching_b200 import synth


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    # a bumpy, asymmetric closed surface (all 6 degrees of freedom observable)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = 0.08 * (1 + 0.3 * np.sin(3 * d[:, 0]) * np.cos(2 * d[:, 1]) + 0.2 * d[:, 2] ** 3)
    pos = (d * r[:, None] * np.array([1.0, 0.7, 0.5])).astype(np.float32)
    # normals: gradient-free stand-in good enough for point-to-plane (unit radial direction, skewed)
    nrm = d / np.array([1.0, 0.7, 0.5]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return pos, nrm.astype(np.float32)


def _numpy_icp(src, tgt, tn, iters, max_dist):
    """binary64 numpy restatement (not bit-exact: different summation order)"""
    s = src.astype(np.float64).copy()
    T = np.eye(4)
    for _ in range(iters):
        d2 = ((s[:, None, :] - tgt[None].astype(np.float64)) ** 2).sum(-1)
        j = d2.argmin(1)
        keep = d2[np.arange(len(s)), j] <= max_dist ** 2
        if keep.sum() < 3:
            break
        ss, dd, nn = s[keep], tgt[j[keep]].astype(np.float64), tn[j[keep]].astype(np.float64)
        A = np.concatenate([np.cross(ss, nn), nn], 1)
        b = (nn * (dd - ss)).sum(1)
        x = np.linalg.solve(A.T @ A, A.T @ b)
        ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
        R = np.array([[cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca],
                      [sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca],
                      [-sb, cb * sa, cb * ca]])
        M = np.eye(4); M[:3, :3] = R; M[:3, 3] = x[3:]
        s = s @ R.T + x[3:]
        T = M @ T
    return T, s


def test_oracle_icp_matches_numpy_restatement():
    tgt, tn = _cloud(700, 1)
    R = synth.axis_angle(np.array([0.3, -0.5, 0.8]), 0.06)
    src = (tgt[::2] @ R.T + np.array([0.004, -0.003, 0.002])).astype(np.float32)
    T, moved, pairs, done, conv = oracle.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    Tn, sn = _numpy_icp(src, tgt, tn, 5, 0.035)
    assert conv and done == 5 and (pairs == len(src)).all()
    assert np.abs(T - Tn).max() < 5e-6
    assert np.abs(moved - sn).max() < 5e-6
    # final = product of the steps: moving the original source by T reproduces the moved cloud
    again = src.astype(np.float64) @ T[:3, :3].astype(np.float64).T + T[:3, 3]
    assert np.abs(again - moved).max() < 2e-6


def test_oracle_icp_recovers_a_small_offset():
    tgt, tn = _cloud(1500, 2)
    R = synth.axis_angle(np.array([1.0, 0.2, -0.4]), 0.03)
    t = np.array([0.002, 0.001, -0.0015])
    src = (tgt @ R.T + t).astype(np.float32)
    before = np.linalg.norm(src - tgt, axis=1).mean()
    T, moved, pairs, done, conv = oracle.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    after = np.linalg.norm(moved - tgt, axis=1).mean()
    assert conv and after < 0.35 * before


def test_oracle_icp_not_converged_below_three_pairs():
    tgt, tn = _cloud(300, 3)
    src = tgt[:50] + np.float32(1.0)          # nothing within 0.035
    T, moved, pairs, done, conv = oracle.icp_point_to_plane(src, tgt, tn, 5, 0.035)
    assert not conv and done == 0 and pairs[0] == 0
    assert np.array_equal(T, np.eye(4, dtype=np.float32))
    assert np.array_equal(moved, src)
    two = tgt[:2].copy()                        # exactly two pairs: still below PCL's minimum of 3
    _, _, pairs, done, conv = oracle.icp_point_to_plane(two, tgt, tn, 5, 0.035)
    assert not conv and pairs[0] == 2 and done == 0


def test_oracle_icp_distance_gate_is_inclusive():
    tgt = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 2]], np.float32)
    tn = np.tile(np.array([0, 0, 1], np.float32), (4, 1))
    # exactly 0.25 away (kept: PCL skips only distance > max), a hair further, far away
    src = np.array([[0, 0, 0.25], [2, 0, 0.25], [0, 2.2500002, 0], [0, 0, 3]], np.float32)
    _, _, pairs, _, _ = oracle.icp_point_to_plane(src, tgt, tn, 1, 0.25)
    assert pairs[0] == 2
