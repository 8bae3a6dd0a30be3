"""Shared small test workloads (seeded, numpy only)."""
import numpy as np

from model_matching_b200 import synth


def object_scene(n_points=3000, n_model=160, radius=0.05, seed=3):
    """A small cluttered scene holding one instance of a bowl model; class probability is high on
    the instance and low on the clutter, as a segmentation network would produce."""
    sc = synth.make_scene(n_points=n_points, extent=(0.40, 0.30, 0.30), n_objects=2, seed=seed,
                          model_radius=radius, model_spacing_pts=4 * n_model)
    mpos, mnrm = synth.make_model(n_model, radius)
    rng = np.random.default_rng(seed)
    is_obj = np.zeros(n_points, bool)
    is_obj[-4 * n_model:] = True
    cls = np.where(is_obj, rng.integers(5000, 10001, n_points), rng.integers(0, 3000, n_points))
    sc["cls"] = (cls.astype(np.float32) * np.float32(1.0 / 10000)).astype(np.float32)
    # a file model is not centred at the origin
    mpos = (mpos + np.array([0.02, -0.01, 0.03], np.float32)).astype(np.float32)
    return sc, mpos, mnrm
