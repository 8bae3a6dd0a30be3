"""Host restatement of clustering::greedy_clustering / get_pose_diff (reference
src/pose_clustering.cpp:5-122), checked against an independent numpy version (CPU only)."""
import os
import subprocess

import numpy as np
import pytest

from model_matching_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "model_matching_b200", "host", "test_clustering")


def _euler_from_R(R):
    # quaternion (Shoemake) -> roll/pitch/yaw as the reference's quaternion_to_euler
    from scipy.spatial.transform import Rotation
    x, y, z, w = Rotation.from_matrix(R).as_quat()
    if w < 0:
        x, y, z, w = -x, -y, -z, -w
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    sp = 2 * (w * y - z * x)
    pitch = np.copysign(np.pi / 2, sp) if abs(sp) >= 1 else np.arcsin(sp)
    yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return np.degrees([roll, pitch, yaw])


def _pose_diff(Tt, Tb, sym):
    e = np.abs(_euler_from_R(np.linalg.inv(Tt[:3, :3]) @ Tb[:3, :3]))
    for d in range(3):
        if sym[d] == 90:
            e[d] = abs(e[d] - 90); e[d] = min(e[d], 90 - e[d])
        elif sym[d] == 180:
            e[d] = min(e[d], 180 - e[d])
        elif sym[d] == 360:
            e[d] = 0
    return e.max(), np.linalg.norm(Tb[:3, 3] - Tt[:3, 3])


def _greedy(T, lcp, frac, best, max_count, min_d, min_a, sym):
    idx = [i for i in range(len(T)) if lcp[i] > np.float32(frac) * np.float32(best)]
    idx.sort(key=lambda i: -lcp[i])
    out = []
    for i in idx:
        if not any((lambda r, t: r < min_a and t < min_d)(*_pose_diff(T[i], T[j], sym)) for j in out):
            out.append(i)
        if len(out) > max_count:
            break
    return out


@pytest.mark.parametrize("sym", [(0, 0, 0), (0, 0, 360), (180, 0, 0)])
def test_greedy_clustering_matches_numpy(sym):
    if not os.path.exists(EXE):
        pytest.skip("host layer not built")
    rng = np.random.Generator(np.random.Philox(5))
    centres = synth.random_rotations(rng, 6)
    T, lcp = [], []
    for c in range(6):
        t0 = rng.uniform(-0.3, 0.3, 3)
        for _ in range(12):
            dR = synth.axis_angle(rng.normal(size=3), np.deg2rad(rng.uniform(0, 6)))
            M = np.eye(4); M[:3, :3] = dR @ centres[c]; M[:3, 3] = t0 + rng.normal(0, 0.004, 3)
            T.append(M); lcp.append(rng.uniform(0.05, 0.6))
    T = np.array(T); lcp = np.array(lcp, np.float32)
    best = float(lcp.max())
    args = dict(frac=0.5, best=best, max_count=10, min_d=0.03, min_a=15.0, sym=sym)
    lines = [f"{len(T)} {args['frac']} {best:.9g} {args['max_count']} {args['min_d']} {args['min_a']} {sym[0]} {sym[1]} {sym[2]}"]
    for M, l in zip(T, lcp):
        lines.append(" ".join(f"{v:.9g}" for v in M.astype(np.float32).T.reshape(16)) + f" {l:.9g}")
    out = subprocess.run([EXE], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    got = [int(x) for x in out if x and not x.startswith("diff")]
    want = _greedy(T.astype(np.float32).astype(np.float64), lcp, **args)
    assert got == want
    assert 2 <= len(got) <= 11
    r, t = [float(x) for x in [l for l in out if l.startswith("diff")][0].split()[1:]]
    wr, wt = _pose_diff(T[0], T[1], sym)
    assert abs(r - wr) < 1e-2 and abs(t - wt) < 1e-5
