"""Drop-in check of the two command lines (reference src/model_preprocess.cpp,
src/stocs_match_one_object.cpp): same argv, same output files, on the reference's YCB example
(BASELINE.json configs[0]) and the instance-mode `packed` example (configs[2])."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "model_matching_b200", "host")


def _run(tmp, scene, obj, env_extra):
    repo = os.path.join(tmp, "repo")
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "examples", scene), os.path.join(repo, "examples", scene))
    os.makedirs(os.path.join(repo, "models", obj))
    shutil.copy(os.path.join(ROOT, "tests", "golden", "models", obj, "textured_vertices.ply"), os.path.join(repo, "models", obj))
    env = dict(os.environ, STOCS_REPO_PATH=repo, STOCS_SEED="7", **env_extra)
    for exe in ("model_preprocess", "stocs_single"):
        assert os.path.exists(os.path.join(HOST, exe)), "build the host layer first (__graft_entry__.build())"
    p1 = subprocess.run([os.path.join(HOST, "model_preprocess"), obj], env=env, capture_output=True, text=True, timeout=300)
    assert p1.returncode == 0, p1.stdout + p1.stderr
    assert os.path.getsize(os.path.join(repo, "models", obj, "ppf_map")) > 1000
    assert os.path.exists(os.path.join(repo, "models", obj, "model_search.ply"))
    scene_dir = os.path.join(repo, "examples", scene)
    p2 = subprocess.run([os.path.join(HOST, "stocs_single"), scene_dir, obj], env=env, capture_output=True, text=True, timeout=300)
    assert p2.returncode == 0, p2.stdout + p2.stderr
    return p1.stdout, p2.stdout, scene_dir


def _check_outputs(scene_dir, obj, out):
    pose = np.loadtxt(os.path.join(scene_dir, f"best_pose_candidate_{obj}.txt"))
    assert pose.shape == (12,)                                   # rows 0-2 of the 4x4, row-major
    R = pose.reshape(3, 4)[:, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-4) and np.isclose(np.linalg.det(R), 1, atol=1e-4)
    for f in ("sampled_scene.ply", "best_pose.ply", "scene.ply"):
        head = open(os.path.join(scene_dir, "dbg", f)).read(200)
        assert head.startswith("ply\nformat ascii 1.0"), f
    assert "Transforms to verify:" in out and "best index:" in out and "|S|:" in out


def test_ycb_class_mode(tmp_path):
    pre, out, scene_dir = _run(str(tmp_path), "ycb", "024_bowl", {})
    assert "After sampling |M|= 472" in pre
    assert "|M| = 472,  |map(M)| = 154951" in out                 # == the oracle's expanded std::map size
    _check_outputs(scene_dir, "024_bowl", out)
    # same seed => same pose (the reference is irreproducible)
    pose1 = open(os.path.join(scene_dir, "best_pose_candidate_024_bowl.txt")).read()
    env = dict(os.environ, STOCS_REPO_PATH=os.path.join(str(tmp_path), "repo"), STOCS_SEED="7")
    subprocess.run([os.path.join(HOST, "stocs_single"), scene_dir, "024_bowl"], env=env, capture_output=True, timeout=300, check=True)
    assert open(os.path.join(scene_dir, "best_pose_candidate_024_bowl.txt")).read() == pose1


def test_packed_instance_mode(tmp_path):
    env = {"STOCS_CAM_INTRINSICS": "615.957763671875,308.1098937988281,615.9578247070312,246.33352661132812",
           "STOCS_DEPTH_SCALE": "0.000125", "STOCS_MODEL_VOXEL_SIZE": "0.005"}
    pre, out, scene_dir = _run(str(tmp_path), "packed", "dove", env)
    assert "After sampling |M|= 944" in pre
    _check_outputs(scene_dir, "dove", out)


def test_linemod_class_mode(tmp_path):
    """LINEMOD settings of the reference README (millimetre model, 1 mm depth units)"""
    env = {"STOCS_CAM_INTRINSICS": "572.4114,325.2611,573.57043,242.04899", "STOCS_DEPTH_SCALE": "0.001",
           "STOCS_MODEL_VOXEL_SIZE": "10", "STOCS_NORMAL_RADIUS": "5", "STOCS_MODEL_SCALE": "0.001"}
    pre, out, scene_dir = _run(str(tmp_path), "linemod", "obj_06", env)
    assert "After sampling |M|= 446" in pre
    assert "|map(M)| = 443361" in out
    _check_outputs(scene_dir, "obj_06", out)
