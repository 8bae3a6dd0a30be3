"""ctypes binding of the CPU oracle (oracle/stocs_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py.  The product package (model_matching_b200) never imports this module.
PARITY UNPINNED: see the header of stocs_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_DIR, "liboracle.so")
    src = os.path.join(_DIR, "stocs_oracle.cpp")
    hdr = os.path.join(_DIR, "..", "model_matching_b200", "csrc", "stocs_math.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    # STOCS_ORACLE_LIB: an arithmetic-model variant built by oracle/sensitivity.py (never the default)
    L = C.CDLL(os.environ.get("STOCS_ORACLE_LIB") or build())
    L.orc_backproject.argtypes = [u16p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_float, f32p, C.c_void_p]
    L.orc_build_scene_cloud.restype = C.c_longlong
    L.orc_build_scene_cloud.argtypes = [u16p, C.c_void_p, u16p, C.c_void_p, C.c_int, C.c_int] + [C.c_float] * 7 + \
        [f32p, f32p, f32p, i32p, f32p, f32p, C.c_longlong]
    L.orc_ppf_compute.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, i32p]
    L.orc_math_eval.argtypes = [C.c_int, f32p, f32p, C.c_int, f32p]
    L.orc_philox.argtypes = [C.c_uint32] * 6 + [u32p]
    L.orc_ppfmap_build.restype = C.c_void_p
    L.orc_ppfmap_build.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int]
    L.orc_ppfmap_free.argtypes = [C.c_void_p]
    L.orc_ppfmap_num_keys.restype = C.c_longlong
    L.orc_ppfmap_num_keys.argtypes = [C.c_void_p]
    L.orc_ppfmap_num_entries.restype = C.c_longlong
    L.orc_ppfmap_num_entries.argtypes = [C.c_void_p]
    L.orc_ppfmap_lookup.restype = C.c_longlong
    L.orc_ppfmap_lookup.argtypes = [C.c_void_p, i32p, i32p, C.c_longlong]
    L.orc_ppfmap_keys.argtypes = [C.c_void_p, i32p]
    L.orc_est_create.restype = C.c_void_p
    L.orc_est_create.argtypes = [f32p, f32p, f32p, C.c_void_p, C.c_int, f32p, f32p, C.c_int,
                                 C.c_void_p, C.c_float, C.c_int, C.c_int]
    L.orc_est_free.argtypes = [C.c_void_p]
    L.orc_est_centroids.argtypes = [C.c_void_p, f32p, f32p]
    L.orc_est_centred.argtypes = [C.c_void_p, f32p, f32p]
    L.orc_est_kd_query.argtypes = [C.c_void_p, f32p, C.c_int, C.c_float, i32p]
    L.orc_est_kd_num_nodes.argtypes = [C.c_void_p]
    L.orc_est_kd_order.argtypes = [C.c_void_p, i32p]
    L.orc_est_score.argtypes = [C.c_void_p, f32p, C.c_longlong, f32p, i32p, C.c_int]
    L.orc_est_score_counters.argtypes = [C.c_void_p, f32p, C.c_longlong, C.c_int, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
    L.orc_icp_point_to_plane.restype = C.c_int
    L.orc_icp_point_to_plane.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p, i32p,
                                         C.POINTER(C.c_int)]
    L.orc_best.argtypes = [f32p, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_float)]
    L.orc_est_sample_class_base.argtypes = [C.c_void_p, C.c_ulonglong, C.c_uint, i32p, f32p,
                                            C.POINTER(C.c_int)]
    L.orc_est_current_prob.argtypes = [C.c_void_p, f32p]
    L.orc_est_class_prob.argtypes = [C.c_void_p, f32p]
    L.orc_est_set_edge_map.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int]
    L.orc_est_sample_instance_base.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_float, i32p, f32p,
                                               C.c_void_p, C.POINTER(C.c_int)]
    L.orc_est_find_congruent.restype = C.c_longlong
    L.orc_est_find_congruent.argtypes = [C.c_void_p, i32p, C.c_float, C.c_float, i32p,
                                         C.c_longlong, i32p]
    L.orc_est_fit.argtypes = [C.c_void_p, i32p, i32p, f32p, f32p]
    L.orc_try_sampled_base.argtypes = [C.c_void_p, i32p, f32p, C.POINTER(C.c_int)]
    _LIB = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def backproject(depth, bgr, fx, cx, fy, cy, depth_scale):
    H, W = depth.shape
    xyz = np.empty((H * W, 3), np.float32)
    rgb = np.empty(H * W, np.uint32)
    d = np.ascontiguousarray(depth, np.uint16)
    b = np.ascontiguousarray(bgr, np.uint8) if bgr is not None else None
    lib().orc_backproject(d, b.ctypes.data if b is not None else None, W, H, fx, cx, fy, cy,
                          depth_scale, xyz, rgb.ctypes.data)
    return xyz, rgb


def build_scene_cloud(depth, bgr, prob, edge, K, depth_scale, voxel_size, class_threshold):
    """CPU restatement of rgbd::load_rgbd_data_sampled's body (src/rgbd.cpp:190-279)."""
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    bgr = None if bgr is None else np.ascontiguousarray(bgr, np.uint8)
    prob = np.ascontiguousarray(prob, np.uint16)
    edge = None if edge is None else np.ascontiguousarray(edge, np.uint8)
    cap = H * W
    pos, nrm, rgb = (np.empty((cap, 3), np.float32) for _ in range(3))
    pix = np.empty((cap, 2), np.int32)
    cls, ep = np.empty(cap, np.float32), np.empty(cap, np.float32)
    k = lib().orc_build_scene_cloud(depth, bgr.ctypes.data if bgr is not None else None, prob,
                                    edge.ctypes.data if edge is not None else None, W, H, K[0], K[1], K[2], K[3],
                                    depth_scale, voxel_size, class_threshold, pos, nrm, rgb, pix, cls, ep, cap)
    return dict(pos=pos[:k].copy(), nrm=nrm[:k].copy(), rgb=rgb[:k].copy(), pix=pix[:k].copy(), cls=cls[:k].copy(),
                edge=ep[:k].copy())


def ppf_compute(p1, n1, p2, n2, tr=5, rot=5):
    p1, n1, p2, n2 = map(_f32, (p1, n1, p2, n2))
    n = p1.reshape(-1, 3).shape[0]
    out = np.empty((n, 4), np.int32)
    lib().orc_ppf_compute(p1, n1, p2, n2, n, tr, rot, out)
    return out


MATH_FN = {"acos": 0, "atan2": 1, "atan": 2, "sin": 3, "cos": 4, "log2": 5}


def math_eval(fn, x, y=None):
    x = _f32(x)
    y = _f32(y) if y is not None else np.zeros_like(x)
    out = np.empty_like(x)
    lib().orc_math_eval(MATH_FN[fn], x, y, x.size, out)
    return out


def philox(ctr, key):
    out = np.empty(4, np.uint32)
    lib().orc_philox(*[int(c) for c in ctr], *[int(k) for k in key], out)
    return out


class PPFMap:
    """The reference's fully expanded std::map (include/rgbd.hpp:23)."""

    def __init__(self, mpos, mnrm, tr=5, rot=5):
        self.mpos, self.mnrm = _f32(mpos), _f32(mnrm)
        self.h = lib().orc_ppfmap_build(self.mpos, self.mnrm, self.mpos.shape[0], tr, rot)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:  # module globals may be gone at interpreter exit
            lib().orc_ppfmap_free(self.h)
            self.h = None

    @property
    def num_keys(self):
        return lib().orc_ppfmap_num_keys(self.h)

    @property
    def num_entries(self):
        return lib().orc_ppfmap_num_entries(self.h)

    def keys(self):
        k = np.empty((self.num_keys, 4), np.int32)
        lib().orc_ppfmap_keys(self.h, k)
        return k

    def lookup(self, key):
        key = np.ascontiguousarray(key, np.int32)
        n = lib().orc_ppfmap_lookup(self.h, key, np.empty((1, 2), np.int32), 0)
        if n < 0:
            return None
        out = np.empty((n, 2), np.int32)
        lib().orc_ppfmap_lookup(self.h, key, out, n)
        return out


class Estimator:
    """Restatement of stocs::stocs_estimator's online methods (src/stocs.cpp)."""

    def __init__(self, spos, snrm, scls, mpos, mnrm, ppfmap=None, spix=None,
                 distance_threshold=0.005, tr=5, rot=5):
        self.S, self.M = len(spos), len(mpos)
        self._keep = (_f32(spos), _f32(snrm), _f32(scls), _f32(mpos), _f32(mnrm), ppfmap)
        pix = np.ascontiguousarray(spix, np.int32) if spix is not None else None
        self._pix = pix
        self.h = lib().orc_est_create(self._keep[0], self._keep[1], self._keep[2],
                                      pix.ctypes.data if pix is not None else None, self.S,
                                      self._keep[3], self._keep[4], self.M,
                                      ppfmap.h if ppfmap is not None else None,
                                      distance_threshold, tr, rot)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:  # module globals may be gone at interpreter exit
            lib().orc_est_free(self.h)
            self.h = None

    def centroids(self):
        a, b = np.empty(3, np.float32), np.empty(3, np.float32)
        lib().orc_est_centroids(self.h, a, b)
        return a, b

    def centred(self):
        s, m = np.empty((self.S, 3), np.float32), np.empty((self.M, 3), np.float32)
        lib().orc_est_centred(self.h, s, m)
        return s, m

    def kd_order(self):
        """original index of every scene point in the reference kd-tree's leaf order, and the node count"""
        out = np.empty(self.S, np.int32)
        lib().orc_est_kd_order(self.h, out)
        return out, int(lib().orc_est_kd_num_nodes(self.h))

    def kd_query(self, q, sqdist):
        q = _f32(q).reshape(-1, 3)
        out = np.empty(q.shape[0], np.int32)
        lib().orc_est_kd_query(self.h, q, q.shape[0], sqdist, out)
        return out

    def score(self, T, threads=1):
        T = _f32(T).reshape(-1, 16)
        H = T.shape[0]
        lcp, inl = np.empty(H, np.float32), np.empty(H, np.int32)
        lib().orc_est_score(self.h, T, H, lcp, inl, threads)
        return lcp, inl

    def score_counters(self, T, threads=1):
        """Data-dependent work of the reference's own loop: dict(queries, examined, hits,
        inliers, bytes) with bytes = 32*queries + 16*examined + 16*hits."""
        T = _f32(T).reshape(-1, 16)
        c = np.zeros(4, np.int64)
        lib().orc_est_score_counters(self.h, T, T.shape[0], threads, c)
        return dict(queries=int(c[0]), examined=int(c[1]), hits=int(c[2]), inliers=int(c[3]),
                    bytes=int(32 * c[0] + 16 * c[1] + 16 * c[2]))

    def sample_class_base(self, seed, base_no):
        ids, inv, st = np.empty(4, np.int32), np.empty(2, np.float32), C.c_int(0)
        ok = lib().orc_est_sample_class_base(self.h, seed, base_no, ids, inv, C.byref(st))
        return bool(ok), ids, inv, st.value

    def set_edge_map(self, edge):
        edge = np.ascontiguousarray(edge, np.uint8)
        self._edge_shape = edge.shape
        lib().orc_est_set_edge_map(self.h, edge, edge.shape[1], edge.shape[0])

    def sample_instance_base(self, seed, base_num, dispersion=0.9):
        ids, inv, st = np.empty(4, np.int32), np.empty(2, np.float32), C.c_int(0)
        mask = np.zeros(self._edge_shape, np.uint8)
        ok = lib().orc_est_sample_instance_base(self.h, seed, base_num, dispersion, ids, inv, mask.ctypes.data, C.byref(st))
        return bool(ok), ids, inv, st.value, mask

    def class_prob(self):
        out = np.empty(self.S, np.float32)
        lib().orc_est_class_prob(self.h, out)
        return out

    def current_prob(self):
        out = np.empty(self.S, np.float32)
        lib().orc_est_current_prob(self.h, out)
        return out

    def find_congruent(self, base, inv1, inv2, cap=1 << 22):
        base = np.ascontiguousarray(base, np.int32)
        q = np.empty((cap, 4), np.int32)
        npq = np.zeros(2, np.int32)
        n = lib().orc_est_find_congruent(self.h, base, float(inv1), float(inv2), q, cap, npq)
        assert n <= cap
        return q[:n].copy(), int(npq[0]), int(npq[1])

    def fit(self, base, quad):
        base = np.ascontiguousarray(base, np.int32)
        quad = np.ascontiguousarray(quad, np.int32)
        Tc, Tw = np.empty(16, np.float32), np.empty(16, np.float32)
        ok = lib().orc_est_fit(self.h, base, quad, Tc, Tw)
        return bool(ok), Tc, Tw

    def try_sampled_base(self, ids):
        ids = np.ascontiguousarray(ids, np.int32).copy()
        inv, ok = np.empty(2, np.float32), C.c_int(0)
        lib().orc_try_sampled_base(self.h, ids, inv, C.byref(ok))
        return bool(ok.value), ids, inv


def best(lcp):
    lcp = _f32(lcp)
    bi, bl = C.c_longlong(0), C.c_float(0)
    lib().orc_best(lcp, lcp.size, C.byref(bi), C.byref(bl))
    return bi.value, bl.value


def icp_point_to_plane(src, tgt, tgt_nrm, max_iterations=5, max_dist=0.035):
    """-> (T 4x4, aligned source, pairs per iteration, iterations done, converged)"""
    src, tgt, tgt_nrm = (_f32(a).reshape(-1, 3) for a in (src, tgt, tgt_nrm))
    T = np.zeros(16, np.float32)
    out = np.empty_like(src)
    pairs = np.zeros(max_iterations, np.int32)
    done = C.c_int(0)
    ok = lib().orc_icp_point_to_plane(src, src.shape[0], tgt, tgt_nrm, tgt.shape[0], max_iterations, max_dist, T, out,
                                      pairs, C.byref(done))
    return T.reshape(4, 4).T.copy(), out, pairs, done.value, bool(ok)
