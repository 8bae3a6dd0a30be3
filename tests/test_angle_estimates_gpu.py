"""The samplers take the integer part of int(atan2(y, x) * 180 / M_PI) -- all a PPF key keeps of an angle --
and the side of 30 degrees an internal angle lies on from fp32 estimates (CUDA atan2f / acosf) whenever
the estimate is farther than 1e-3 degree from an integer (1e-2 from 30 / 150), and from the pinned
binary64 evaluation otherwise (csrc/ppf_device.cuh).  That is exact as long as the estimate stays within
the guard band of the pinned value: measured here on random and on adversarial inputs, together with the
decisions themselves."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs():
    rng = np.random.default_rng(11)
    n = 1 << 20
    th = rng.uniform(0, np.pi, n)
    r = 10.0 ** rng.uniform(-6, 2, n)
    y, x = [np.abs(r * np.sin(th))], [r * np.cos(th)]
    # adversarial: angles a hair's breadth from every integer degree, on both sides
    k = np.arange(0, 181, dtype=np.float64)
    for d in (0.0, 1e-7, 1e-6, 1e-5, 1e-4, 5e-4, 9e-4, 1.1e-3, 1e-2):
        for s in (-1, 1):
            a = np.clip(np.deg2rad(k + s * d), 0, np.pi)
            y.append(np.abs(np.sin(a))); x.append(np.cos(a))
    # axes and degenerate pairs
    y.append(np.array([0, 0, 0, 1, 1e-30, 0, np.nan, 1, np.inf], np.float64))
    x.append(np.array([1, -1, 0, 0, 1, 1e-30, 1, np.nan, 1], np.float64))
    return np.concatenate(y).astype(np.float32), np.concatenate(x).astype(np.float32)


def test_fp32_angle_estimates_stay_inside_their_guard_band(gpu_ctx):
    y, x = _inputs()
    r = gpu_ctx.debug_angle_estimates(y, x)
    ok = np.isfinite(r["pinned"]) & np.isfinite(r["est"])
    err = np.abs(r["est"][ok].astype(np.float64) - r["pinned"][ok])
    assert ok.sum() > 1_000_000
    assert err.max() < 2e-4, err.max()            # guard band: 1e-3
    assert np.array_equal(r["fast_floor"], r["pinned_floor"])
    # the fast path really is the common one: the pinned integer agrees with the estimate's floor almost always
    assert (np.floor(r["est"][ok]) == r["pinned_floor"][ok]).mean() > 0.99


def test_thirty_degree_predicate_decisions(gpu_ctx):
    rng = np.random.default_rng(12)
    d = [rng.uniform(-1.0, 1.0, 1 << 20)]
    for c in (30.0, 150.0):
        for off in (0.0, 1e-6, 1e-5, 1e-4, 1e-3, 5e-3, 9e-3, 1.1e-2, 0.1):
            for s in (-1, 1):
                d.append(np.cos(np.deg2rad(c + s * off)) + np.arange(-8, 9) * 6e-8)
    d.append(np.array([-1.0, 1.0, 0.0, 1.0000001, -1.0000001, np.nan, 2.0]))
    d = np.concatenate(d).astype(np.float32)
    r = gpu_ctx.debug_angle_estimates(np.zeros_like(d), d)
    assert np.array_equal(r["below30_fast"], r["below30_pinned"])
    assert 0.1 < r["below30_pinned"].mean() < 0.2      # d uniform: 2 * (1 - cos 30) / 2 = 0.134
