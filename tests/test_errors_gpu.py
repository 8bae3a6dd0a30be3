"""Error convention of the C ABI (status codes + stocs_b200_last_error), call-order checks and
argument validation: the library must fail loudly, never fall back."""
import ctypes as C

import numpy as np
import pytest

import model_matching_b200 as mm
from model_matching_b200 import Context, StocsError, synth

pytestmark = pytest.mark.gpu


def test_call_order_and_arguments():
    ctx = Context(0)
    L = mm.lib()
    T = np.eye(4, dtype=np.float32).T.reshape(1, 16).copy()
    with pytest.raises(StocsError, match="upload_model and upload_scene first"):
        ctx.score_lcp(T)
    with pytest.raises(StocsError, match="upload_model"):
        ctx.sample_bases(1, 0, 4)
    mpos, mnrm = synth.make_model(64)
    ctx.upload_model(mpos, mnrm)
    with pytest.raises(StocsError):
        ctx.score_lcp(T)                                   # still no scene
    sc = synth.make_scene(n_points=5000, extent=(0.3, 0.3, 0.3), n_objects=1, seed=2, model_radius=0.04, model_spacing_pts=256)
    ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    lcp, inl = ctx.score_lcp(T)
    assert lcp.shape == (1,)
    # parameters are frozen once data is uploaded
    assert L.stocs_b200_set_params(ctx.h, 0.01, 5, 5) == -3
    assert b"set_params" in L.stocs_b200_last_error(ctx.h)
    # bad arguments
    assert L.stocs_b200_score_lcp(ctx.h, None, 5, None, None) == -1
    assert L.stocs_b200_upload_model(ctx.h, None, None, 10) == -1
    big = np.zeros((6000, 3), np.float32)
    with pytest.raises(StocsError, match="at most"):
        ctx.upload_model(big, big)
    with pytest.raises(StocsError, match="reduce_best"):
        ctx.reduce_best(np.ones(10, np.float32), K=64)
    bad = sc["pos"].copy(); bad[7, 1] = np.nan
    with pytest.raises(StocsError, match="non-finite"):
        ctx.upload_scene(bad, sc["nrm"], sc["cls"])
    # instance sampling needs an edge map and pixel coordinates
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    with pytest.raises(StocsError, match="upload_edge_map"):
        ctx.sample_instance_base(1, 1)
    ctx.upload_edge_map(np.full((480, 640), 255, np.uint8))
    with pytest.raises(StocsError, match="pixel"):
        ctx.sample_instance_base(1, 1)
    # bad device ordinal / NULL handle
    h = C.c_void_p()
    assert L.stocs_b200_create(C.byref(h), 9999) < 0
    assert L.stocs_b200_get_counters(None, None, 0) == -1
    ctx.close()


def test_two_contexts_are_independent():
    a, b = Context(0), Context(0, distance_threshold=0.01)
    mpos, mnrm = synth.make_model(96)
    sc = synth.make_scene(n_points=8000, extent=(0.4, 0.3, 0.3), n_objects=2, seed=5, model_radius=0.05, model_spacing_pts=512)
    T, _ = synth.make_hypotheses(400, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=1, near_fraction=0.2)
    for c in (a, b):
        c.upload_model(mpos, mnrm)
        c.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
    la, ia = a.score_lcp(T)
    lb, ib = b.score_lcp(T)
    assert np.all(ib >= ia) and ib.sum() > ia.sum()        # a larger radius can only add inliers
    la2, ia2 = a.score_lcp(T)
    assert np.array_equal(la, la2) and np.array_equal(ia, ia2)
    a.close(); b.close()
