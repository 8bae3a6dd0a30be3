"""Small end-to-end run used under compute-sanitizer (memcheck / racecheck): every kernel of the
library on tiny inputs, checked against the oracle.  Not collected by pytest."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from model_matching_b200 import Context, synth  # noqa: E402
from scenes import object_scene  # noqa: E402

sc, mpos, mnrm = object_scene(n_points=1500, n_model=64, radius=0.04)
pix = np.stack([np.arange(1500) % 480, np.arange(1500) % 640], -1).astype(np.int32)
edge = np.full((480, 640), 255, np.uint8); edge[::30] = 0
ctx = Context(0)
ctx.upload_model(mpos, mnrm)
ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"], pix)
T, _ = synth.make_hypotheses(96, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=1, near_fraction=0.3)
lcp, inl = ctx.score_lcp(T)
est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
ol, oi = est.score(T)
assert np.array_equal(inl, oi) and np.array_equal(lcp.view(np.uint32), ol.view(np.uint32))
ctx.reduce_best(lcp, K=8)
ctx.select_above(lcp, 0.01)
ids, inv, ok = ctx.sample_bases(3, 0, 6)
if ok.any():
    q, off = ctx.find_congruent(ids[ok], inv[ok])
    if len(q):
        ctx.fit_transforms(np.repeat(ids[ok][:1], min(len(q), 8), 0), q[:8])
ctx.run_pipeline(5, 6, 10)
ctx.upload_edge_map(edge)
for b in range(1, 4):
    ctx.sample_instance_base(3, b)
depth = (1000 + np.arange(64 * 48).reshape(48, 64) % 37).astype(np.uint16)
ctx.backproject(depth, None, 50.0, 32.0, 50.0, 24.0, 0.001)
ctx.build_scene_cloud(depth, None, np.full((48, 64), 9000, np.uint16), None, [50.0, 32.0, 50.0, 24.0], 0.001, 0.005, 0.1)
ctx.ppf_lookup(np.array([20, 30, 30, 30], np.int32))
ctx.close()
print("sanitize_smoke ok")
