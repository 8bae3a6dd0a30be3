"""GPU parity of the scene-cloud construction (reference src/rgbd.cpp:190-279: back-projection,
VoxelGrid, RadiusOutlierRemoval, re-projection, class threshold, depth normals) against the CPU
oracle on the reference's three example frames (configs[0..2]) and on a synthetic frame."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "examples")

CASES = {
    "ycb": ("024_bowl", [1066.778, 312.986, 1067.487, 241.310], 1 / 10000.0, False),
    "linemod": ("obj_06", [572.4114, 325.2611, 573.57043, 242.04899], 1 / 1000.0, False),
    "packed": ("dove", [615.957763671875, 308.1098937988281, 615.9578247070312, 246.33352661132812], 1 / 8000.0, True),
}


def _read(scene, obj, with_edge):
    import cv2
    d = os.path.join(G, scene)
    depth = cv2.imread(os.path.join(d, "depth.png"), cv2.IMREAD_UNCHANGED)
    bgr = cv2.imread(os.path.join(d, "rgb.png"), cv2.IMREAD_COLOR)
    prob = cv2.imread(os.path.join(d, "probability_maps", obj + ".png"), cv2.IMREAD_UNCHANGED)
    edge = cv2.imread(os.path.join(d, "probability_maps", "edge.png"), cv2.IMREAD_GRAYSCALE) if with_edge else None
    return depth, bgr, prob, edge


def _same(a, b):
    assert a["pos"].shape == b["pos"].shape
    for k in ("pos", "nrm", "rgb", "pix", "cls", "edge"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("scene", sorted(CASES))
def test_example_frames_bit_exact(gpu_ctx, scene):
    obj, K, scale, with_edge = CASES[scene]
    depth, bgr, prob, edge = _read(scene, obj, with_edge)
    want = oracle.build_scene_cloud(depth, bgr, prob, edge, K, scale, 0.005, 0.10)
    got = gpu_ctx.build_scene_cloud(depth, bgr, prob, edge, K, scale, 0.005, 0.10)
    assert len(want["pos"]) > 1000
    _same(got, want)


def test_synthetic_frame_and_edge_cases(gpu_ctx):
    rng = np.random.default_rng(1234)
    H, W = 480, 640
    yy, xx = np.mgrid[0:H, 0:W]
    depth = (1000 + 0.2 * xx + 0.1 * yy + rng.normal(0, 1.0, (H, W))).astype(np.uint16)   # plane at ~1 m (mm)
    depth[100:200, 100:300] -= 150                                                         # a box in front
    depth[rng.uniform(size=depth.shape) < 0.05] = 0                                        # holes
    depth[:, :40] = 0
    bgr = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    prob = rng.integers(0, 10001, (H, W)).astype(np.uint16)
    edge = rng.integers(0, 256, (H, W)).astype(np.uint8)
    K = [572.4114, 325.2611, 573.57043, 242.04899]
    for voxel, thr in ((0.005, 0.10), (0.01, 0.5)):
        want = oracle.build_scene_cloud(depth, bgr, prob, edge, K, 1 / 1000.0, voxel, thr)
        got = gpu_ctx.build_scene_cloud(depth, bgr, prob, edge, K, 1 / 1000.0, voxel, thr)
        assert len(want["pos"]) > 500
        _same(got, want)
    # everything filtered: all depth zero
    z = np.zeros((H, W), np.uint16)
    got = gpu_ctx.build_scene_cloud(z, None, prob, None, K, 1 / 1000.0, 0.005, 0.1)
    assert len(got["pos"]) == 0
    # beyond 2 m: dropped by the z test
    far = np.full((H, W), 3000, np.uint16)
    got = gpu_ctx.build_scene_cloud(far, None, prob, None, K, 1 / 1000.0, 0.005, 0.0)
    want = oracle.build_scene_cloud(far, None, prob, None, K, 1 / 1000.0, 0.005, 0.0)
    assert len(got["pos"]) == len(want["pos"]) == 0
