"""Pins of the shared leaf arithmetic (model_matching_b200/csrc/stocs_math.h) as compiled by g++:
independent references are numpy's double-precision libm (rounded once to binary32), published
Random123 known-answer vectors, and hand-computed cases of the reference's integer rules."""
import numpy as np
import pytest

import oracle


def _ulp_diff(got, ref64):
    ref = ref64.astype(np.float32)
    ok = ~(np.isnan(got) & np.isnan(ref))
    return np.max(np.abs(got.view(np.int32).astype(np.int64)[ok] - ref.view(np.int32).astype(np.int64)[ok]))


@pytest.mark.parametrize("fn,lo,hi,ref", [
    ("acos", -1.0, 1.0, np.arccos), ("atan", 0.0, 3.2, np.arctan),
    ("sin", 0.0, 6.3, np.sin), ("cos", 0.0, 6.3, np.cos), ("log2", 1e-4, 16.0, np.log2)])
def test_unary_functions_are_correctly_rounded(fn, lo, hi, ref):
    rng = np.random.default_rng(0)
    x = rng.uniform(lo, hi, 100000).astype(np.float32)
    assert _ulp_diff(oracle.math_eval(fn, x), ref(x.astype(np.float64))) <= 1


def test_atan2_and_special_values():
    rng = np.random.default_rng(1)
    y = np.abs(rng.normal(size=100000)).astype(np.float32)
    x = rng.normal(size=100000).astype(np.float32)
    assert _ulp_diff(oracle.math_eval("atan2", x, y), np.arctan2(y.astype(np.float64), x.astype(np.float64))) <= 1
    sp = oracle.math_eval("atan2", np.array([1, -1, 0, 0], np.float32), np.array([0, 0, 1, 0], np.float32))
    assert np.allclose(sp, [0, np.pi, np.pi / 2, 0])
    ac = oracle.math_eval("acos", np.array([1, -1, 0, 1.0000001, -1.0000001, np.nan], np.float32))
    assert ac[0] == 0 and np.isclose(ac[1], np.pi) and np.isclose(ac[2], np.pi / 2) and np.all(np.isnan(ac[3:]))


def test_angle_predicate_is_monotone_around_30_degrees():
    """The kernels replace acos(d)*180/pi < 30 by a threshold on d found by bisection
    (capi.cu stocs_angle_threshold_dot); that needs the predicate to flip exactly once."""
    lo, hi = np.float32(0.85), np.float32(0.88)
    bits = np.arange(lo.view(np.uint32), hi.view(np.uint32), dtype=np.uint32)
    d = bits.view(np.float32)
    ang = ((oracle.math_eval("acos", d) * np.float32(180.0)).astype(np.float64) / np.pi).astype(np.float32)
    pred = ang < np.float32(30)
    flips = np.flatnonzero(pred[1:] != pred[:-1])
    assert flips.size == 1 and not pred[0] and pred[-1]
    assert abs(float(d[flips[0] + 1]) - np.cos(np.pi / 6)) < 1e-6


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10
    assert [hex(v) for v in oracle.philox((0, 0, 0, 0), (0, 0))] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in oracle.philox((0xffffffff,) * 4, (0xffffffff,) * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in oracle.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_ppf_known_answers():
    """reference src/rgbd.cpp:85-121.  Two points 23.7 mm apart along x, normals +z and +x:
    f1 = int(23.7) = 23 -> nearest multiple of 5 with ties up = 25; angle(n1,u)=90, angle(n2,u)=180
    (u = p1-p2 = -x), angle(n1,n2) = 90."""
    p1, n1 = np.array([[0, 0, 0]], np.float32), np.array([[0, 0, 1]], np.float32)
    p2, n2 = np.array([[0.0237, 0, 0]], np.float32), np.array([[1, 0, 0]], np.float32)
    assert oracle.ppf_compute(p1, n1, p2, n2).tolist() == [[25, 90, 180, 90]]
    # asymmetry (quirk 9): swapping the points changes the second / third feature
    assert oracle.ppf_compute(p2, n2, p1, n1).tolist() == [[25, 0, 90, 90]]
    # truncate to int mm first, then nearest multiple of 5 (strict <: 22 -> 20, 23 -> 25, 27 -> 25, 28 -> 30)
    for dist, want in [(0.0221, 20), (0.0229, 20), (0.0231, 25), (0.02749, 25), (0.0276, 25), (0.0281, 30), (0.0051, 5), (0.0049, 5), (0.0024, 0)]:
        q = np.array([[dist, 0, 0]], np.float32)
        assert oracle.ppf_compute(p1, n1, q, n2)[0, 0] == want, dist


def test_backprojection_formula():
    depth = np.array([[0, 1000], [2000, 65535]], np.uint16)
    bgr = np.array([[[1, 2, 3], [4, 5, 6]], [[7, 8, 9], [255, 0, 128]]], np.uint8)
    fx, cx, fy, cy = 572.4114, 325.2611, 573.57043, 242.04899
    xyz, rgb = oracle.backproject(depth, bgr, fx, cx, fy, cy, 1 / 1000.0)
    f = np.float32
    d = depth.astype(np.float32).ravel() * f(1 / 1000.0)
    j = np.array([0, 1, 0, 1], np.float32); i = np.array([0, 0, 1, 1], np.float32)
    assert np.array_equal(xyz[:, 0], (j - f(cx)) * d / f(fx))
    assert np.array_equal(xyz[:, 1], (i - f(cy)) * d / f(fy))
    assert np.array_equal(xyz[:, 2], d)
    assert rgb.tolist() == [0x030201, 0x060504, 0x090807, 0x8000ff]
