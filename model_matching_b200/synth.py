"""Synthetic scene / model / hypothesis generators for the StoCS hot path.

These produce the inputs named in BASELINE.json configs[3] ("synthetic 640x480 depth scene,
1M-point cloud, 1e6 hypotheses") and SURVEY.md section 8(d) "S1": surfaces sampled on a 5 mm
lattice jittered by +-1 mm with outward unit normals and class probability U{0..10000}/10000,
a 512-point bowl-like model at 10 mm spacing, and a hypothesis mix of 1 % near-truth poses and
99 % uniform SO(3) x scene-AABB poses.  Everything is seeded (numpy Philox) and pure numpy;
nothing here touches the GPU or the oracle.
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def _plane_patch(origin, eu, ev, lu, lv, normal, spacing):
    """Lattice points on the parallelogram origin + a*eu + b*ev, a<lu, b<lv."""
    nu, nv = max(int(round(lu / spacing)), 1), max(int(round(lv / spacing)), 1)
    a = (np.arange(nu) + 0.5) * spacing
    b = (np.arange(nv) + 0.5) * spacing
    A, B = np.meshgrid(a, b, indexing="ij")
    p = origin[None, :] + A.reshape(-1, 1) * eu[None, :] + B.reshape(-1, 1) * ev[None, :]
    n = np.broadcast_to(np.asarray(normal, np.float64), p.shape)
    return p, n


def _box(center, size, spacing):
    """Five visible faces (no bottom) of an axis-aligned box resting on z = center.z - size.z/2."""
    cx, cy, cz = center
    sx, sy, sz = size
    ex, ey, ez = np.eye(3)
    lo = np.array([cx - sx / 2, cy - sy / 2, cz - sz / 2])
    P, N = [], []
    for (o, eu, ev, lu, lv, n) in [
        (lo + ez * sz, ex, ey, sx, sy, ez),          # top
        (lo, ex, ez, sx, sz, -ey),                   # front  (y = lo)
        (lo + ey * sy, ex, ez, sx, sz, ey),          # back
        (lo, ey, ez, sy, sz, -ex),                   # left
        (lo + ex * sx, ey, ez, sy, sz, ex),          # right
    ]:
        p, nn = _plane_patch(o, eu, ev, lu, lv, n, spacing)
        P.append(p)
        N.append(nn)
    return np.concatenate(P), np.concatenate(N)


def _cylinder(center, radius, height, spacing):
    cx, cy, cz = center
    nth = max(int(round(2 * np.pi * radius / spacing)), 8)
    nz = max(int(round(height / spacing)), 1)
    th = (np.arange(nth) + 0.5) * (2 * np.pi / nth)
    zz = cz - height / 2 + (np.arange(nz) + 0.5) * spacing
    TH, ZZ = np.meshgrid(th, zz, indexing="ij")
    side = np.stack([cx + radius * np.cos(TH), cy + radius * np.sin(TH), ZZ], -1).reshape(-1, 3)
    nside = np.stack([np.cos(TH), np.sin(TH), np.zeros_like(TH)], -1).reshape(-1, 3)
    # top disc
    g = (np.arange(int(2 * radius / spacing) + 1) + 0.5) * spacing - radius
    X, Y = np.meshgrid(g, g, indexing="ij")
    m = X * X + Y * Y <= radius * radius
    top = np.stack([cx + X[m], cy + Y[m], np.full(m.sum(), cz + height / 2)], -1)
    ntop = np.broadcast_to(np.array([0.0, 0.0, 1.0]), top.shape)
    return np.concatenate([side, top]), np.concatenate([nside, ntop])


def bowl_points(n_points, radius=0.09):
    """Fibonacci lattice on the lower hemisphere of a sphere (a bowl opening towards +z),
    outward normals.  n_points=512, radius=0.09 gives ~10 mm spacing on an 18 cm bowl."""
    i = np.arange(n_points) + 0.5
    z = -(i / n_points)                        # uniform in z <=> uniform in area
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = i * (np.pi * (3.0 - np.sqrt(5.0)))
    n = np.stack([r * np.cos(phi), r * np.sin(phi), z], -1)
    return (radius * n).astype(np.float32), n.astype(np.float32)


def make_model(n_points=512, radius=0.09):
    pos, nrm = bowl_points(n_points, radius)
    nrm = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    return pos.astype(np.float32), nrm.astype(np.float32)


def random_rotations(rng, n):
    """Uniform SO(3) from normalised Gaussian quaternions; returns (n,3,3) float64."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def axis_angle(axis, angle):
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    c, s = np.cos(angle), np.sin(angle)
    C = 1 - c
    R = np.stack([np.stack([c + x * x * C, x * y * C - z * s, x * z * C + y * s], -1),
                  np.stack([y * x * C + z * s, c + y * y * C, y * z * C - x * s], -1),
                  np.stack([z * x * C - y * s, z * y * C + x * s, c + z * z * C], -1)], -2)
    return R


def make_scene(n_points=1 << 20, seed=1234, extent=(2.0, 1.5, 1.4), n_objects=64,
               spacing=0.005, jitter=0.001, model_radius=0.09, model_spacing_pts=2048):
    """Returns dict(pos, nrm, cls, gt_R, gt_t): a shelving unit (4 boards + 3 walls) carrying
    n_objects boxes/cylinders and one instance of the bowl model at the ground-truth pose
    (gt_R, gt_t: model frame -> scene frame).  Exactly n_points points."""
    rng = _rng(seed)
    ex, ey, ez = extent
    scale = 1.0
    P, N = [], []
    shelves = [0.0, ez * 0.25, ez * 0.5, ez * 0.75]
    e = np.eye(3)
    for zs in shelves:
        p, n = _plane_patch(np.array([0, 0, zs]), e[0], e[1], ex, ey, e[2], spacing)
        P.append(p); N.append(n)
    p, n = _plane_patch(np.array([0, ey, 0.0]), e[0], e[2], ex, ez, -e[1], spacing); P.append(p); N.append(n)
    p, n = _plane_patch(np.array([0, 0, 0.0]), e[1], e[2], ey, ez, e[0], spacing); P.append(p); N.append(n)
    p, n = _plane_patch(np.array([ex, 0, 0.0]), e[1], e[2], ey, ez, -e[0], spacing); P.append(p); N.append(n)
    # objects, laid out on a jittered grid per shelf so they do not overlap the bowl slot (slot 0)
    per = int(np.ceil((n_objects + 1) / len(shelves)))
    gx = int(np.ceil(np.sqrt(per * ex / ey)))
    gy = int(np.ceil(per / gx))
    slot = 0
    gt_R = axis_angle(np.array([0.3, -0.5, 0.8]), 0.7)
    gt_t = None
    for si, zs in enumerate(shelves):
        for k in range(per):
            if slot > n_objects:
                break
            cx = (k % gx + 0.5) * ex / gx + rng.uniform(-0.02, 0.02)
            cy = (k // gx + 0.5) * ey / gy + rng.uniform(-0.02, 0.02)
            if slot == 0:
                gt_t = np.array([cx, cy, zs + model_radius * 1.2])
            else:
                if rng.uniform() < 0.5:
                    s = rng.uniform(0.08, 0.25, size=3) * scale
                    s[2] = min(s[2], ez * 0.25 - 0.03)
                    p, n = _box((cx, cy, zs + s[2] / 2), s, spacing)
                else:
                    r = rng.uniform(0.03, 0.08) * scale
                    h = min(rng.uniform(0.08, 0.25), ez * 0.25 - 0.03)
                    p, n = _cylinder((cx, cy, zs + h / 2), r, h, spacing)
                P.append(p); N.append(n)
            slot += 1
    P = np.concatenate(P); N = np.concatenate(N)
    P = P + rng.uniform(-jitter, jitter, size=P.shape)
    # bowl instance at 5 mm spacing (4x the model's point count), jittered the same way
    bp, bn = bowl_points(model_spacing_pts, model_radius)
    bp = bp.astype(np.float64) @ gt_R.T + gt_t
    bn = bn.astype(np.float64) @ gt_R.T
    bp = bp + rng.uniform(-jitter, jitter, size=bp.shape)
    n_rest = n_points - bp.shape[0]
    if P.shape[0] >= n_rest:
        sel = np.sort(rng.choice(P.shape[0], size=n_rest, replace=False))
        P, N = P[sel], N[sel]
    else:  # pad by re-sampling jittered copies (tiny test scenes never get here)
        extra = rng.choice(P.shape[0], size=n_rest - P.shape[0], replace=True)
        P = np.concatenate([P, P[extra] + rng.uniform(-jitter, jitter, size=(extra.size, 3))])
        N = np.concatenate([N, N[extra]])
    pos = np.concatenate([P, bp]).astype(np.float32)
    nrm = np.concatenate([N, bn])
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    cls = (rng.integers(0, 10001, size=pos.shape[0]).astype(np.float32) * np.float32(1.0 / 10000)).astype(np.float32)
    return dict(pos=pos, nrm=nrm, cls=cls, gt_R=gt_R, gt_t=gt_t)


def class_map(n_points, seed):
    """Per-object class probability of every scene point, U{0..10000}/10000 (seeds 1000+k)."""
    rng = _rng(seed)
    return (rng.integers(0, 10001, size=n_points).astype(np.float32) * np.float32(1.0 / 10000)).astype(np.float32)


def to_colmajor16(R, t):
    """(n,3,3),(n,3) -> (n,16) float32 in Eigen::Matrix4f memory order."""
    n = R.shape[0]
    T = np.zeros((n, 4, 4), np.float32)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    T[:, 3, 3] = 1
    return np.ascontiguousarray(T.transpose(0, 2, 1).reshape(n, 16))


def make_hypotheses(H, scene_pos, model_pos, gt_R, gt_t, seed=4321, near_fraction=0.01,
                    max_rot_deg=2.0, max_trans=0.002):
    """H transforms in the CENTRED frames the estimator scores in (src/stocs.cpp:943-964):
    q = T (m - c_model) should land on (s - c_scene).  1 % near-truth, 99 % uniform."""
    rng = _rng(seed)
    cs = scene_pos.astype(np.float64).mean(0)
    cm = model_pos.astype(np.float64).mean(0)
    lo = scene_pos.min(0).astype(np.float64) - cs
    hi = scene_pos.max(0).astype(np.float64) - cs
    n_near = int(round(H * near_fraction))
    near_idx = np.sort(rng.choice(H, size=n_near, replace=False)) if n_near else np.zeros(0, np.int64)
    R = random_rotations(rng, H)
    t = rng.uniform(lo, hi, size=(H, 3))
    if n_near:
        ax = rng.normal(size=(n_near, 3))
        ang = np.deg2rad(rng.uniform(0, max_rot_deg, size=n_near))
        dR = axis_angle(ax, ang)
        Rn = dR @ gt_R[None]
        d = rng.normal(size=(n_near, 3))
        d *= (rng.uniform(0, max_trans, size=(n_near, 1)) / np.linalg.norm(d, axis=1, keepdims=True))
        tn = (Rn @ cm) + gt_t[None] - cs[None] + d
        R[near_idx] = Rn
        t[near_idx] = tn
    return to_colmajor16(R, t), near_idx
