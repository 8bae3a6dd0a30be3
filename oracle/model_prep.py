"""numpy restatement of the OFFLINE model preparation that precedes the PPF table
(stocs::pre_process_model, reference src/stocs.cpp:41-60).  TEST INFRASTRUCTURE ONLY.

The reference calls PCL (whose sources are not in its tree; version unpinned): this file restates
the PUBLISHED algorithms of the three operators with plain numpy / scipy so that the C++
restatement in model_matching_b200/host/rgbd.cpp has an independent statement to be checked against
(tests/test_model_prep.py).  "Parity unpinned" with respect to a real PCL build.

 1. pcl::NormalEstimation, radius search r, viewpoint (0,0,0)   (src/rgbd.cpp:72-83)
      neighbours = points within r (the query point included); fewer than 3 -> NaN normal;
      normal = eigenvector of the smallest eigenvalue of the neighbours' covariance matrix,
      flipped so that it points towards the viewpoint: n . (vp - p) >= 0.
 2. normals negated ("normals face outside", src/stocs.cpp:48-52)
 3. pcl::VoxelGrid, leaf L                                        (src/stocs.cpp:54-57)
      leaf index = floor(p / L) - floor(min / L) per axis, linear index x + y*dx + z*dx*dy;
      one output point per occupied leaf, ALL fields (position, normal) averaged, output in
      increasing leaf index.
 4. load_ply_model (src/rgbd.cpp:12-33): points whose normal is not finite are dropped, positions
      scaled, normals normalised.
"""
import numpy as np


def read_ply_xyz(path):
    with open(path) as f:
        n = 0
        for line in f:
            t = line.split()
            if t[:2] == ["element", "vertex"]:
                n = int(t[2])
            if t and t[0] == "end_header":
                break
        return np.loadtxt(f, dtype=np.float64, max_rows=n, usecols=(0, 1, 2)).astype(np.float32)


def estimate_normals(pts, radius):
    from scipy.spatial import cKDTree
    p64 = pts.astype(np.float64)
    tree = cKDTree(p64)
    nrm = np.full((len(pts), 3), np.nan)
    r2 = np.float32(radius) * np.float32(radius)
    for i, nb in enumerate(tree.query_ball_point(p64, float(radius) * (1 + 1e-6))):
        nb = np.asarray(nb)
        d = pts[nb] - pts[i]                                   # binary32 distance test, as PCL's kd-tree search does
        nb = nb[(d * d).sum(1, dtype=np.float32) <= r2]
        if len(nb) < 3:
            continue
        q = p64[nb] - p64[nb].mean(0)
        w, v = np.linalg.eigh(q.T @ q)
        n = v[:, 0]
        if n @ (-p64[i]) < 0:
            n = -n
        nrm[i] = n
    return nrm.astype(np.float32)


def voxel_grid(pts, nrm, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(pts * inv).astype(np.int64)
    ijk -= ijk.min(0)
    dx, dy, _ = ijk.max(0) + 1
    key = ijk[:, 0] + ijk[:, 1] * dx + ijk[:, 2] * dx * dy
    order = np.argsort(key, kind="stable")
    uk, start = np.unique(key[order], return_index=True)
    cnt = np.diff(np.append(start, len(key)))
    P = np.add.reduceat(pts[order].astype(np.float64), start) / cnt[:, None]
    N = np.add.reduceat(nrm[order].astype(np.float64), start) / cnt[:, None]     # NaN propagates, as in PCL
    return P.astype(np.float32), N.astype(np.float32)


def prepare_model(path, normal_radius, scale, voxel_size):
    """-> (positions, unit normals) of models/<obj>/model_search.ply"""
    pts = read_ply_xyz(path)
    nrm = -estimate_normals(pts, normal_radius)
    P, N = voxel_grid(pts, nrm, voxel_size)
    ok = np.isfinite(N).all(1)
    P, N = P[ok] * np.float32(scale), N[ok]
    N = N / np.linalg.norm(N, axis=1, keepdims=True)
    return P.astype(np.float32), N.astype(np.float32)
