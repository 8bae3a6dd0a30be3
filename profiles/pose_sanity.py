"""Pose sanity on the reference's three example scenes (BASELINE.json configs[0..2]); run on the GPU box.

For each scene: stocs_single with 20 different seeds (100 bases, <= 200 sets per base: the
reference's settings) and one long run (5000 bases) as the best pose this method can find on the
frame.  The reference ships no ground-truth pose, so agreement is measured between runs with the
symmetry-aware ADD-S distance (mean over model points of the distance to the nearest model point
under the other pose): how many of the 20 short runs land within 1 cm of the long run's pose, and
how tight the modal cluster of the short runs is.  Writes profiles/r02_pose_sanity.json.

    python profiles/pose_sanity.py [out.json]
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_shim_gpu import SCENES, make_tree  # noqa: E402

HOST = os.path.join(ROOT, "model_matching_b200", "host")


def read_model(path):
    with open(path) as f:
        for line in f:
            if line.startswith("element vertex"):
                n = int(line.split()[2])
            if line.startswith("end_header"):
                break
        return np.loadtxt(f, max_rows=n)[:, :3]


def adds(model, A, B):
    from scipy.spatial import cKDTree
    a = model @ A[:, :3].T + A[:, 3]
    b = model @ B[:, :3].T + B[:, 3]
    return float(cKDTree(b).query(a)[0].mean())


def run(scene, seed, bases, tmp):
    obj, env_scene = SCENES[scene]
    repo = os.path.join(tmp, "repo")
    scene_dir = os.path.join(repo, "examples", scene)
    env = dict(os.environ, STOCS_REPO_PATH=repo, STOCS_SEED=str(seed), STOCS_NUM_BASES=str(bases), **env_scene)
    p = subprocess.run([os.path.join(HOST, "stocs_single"), scene_dir, obj], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    out = p.stdout
    m = re.search(r"best index: (-?\d+), maximum score: ([0-9.e+-]+)", out)
    r = {"seed": seed, "bases_valid": int(re.search(r"Sampled (\d+) bases", out).group(1)),
         "congruent_sets": int(re.search(r"found (\d+) congruent sets", out).group(1)),
         "transforms": int(re.search(r"Transforms to verify: (\d+)", out).group(1)),
         "best_lcp": float(m.group(2)), "pose": None}
    f = os.path.join(scene_dir, f"best_pose_candidate_{obj}.txt")
    if int(m.group(1)) >= 0 and os.path.exists(f):
        r["pose"] = np.loadtxt(f).reshape(3, 4).tolist()
        os.remove(f)
    return r


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_pose_sanity.json")
    report = {}
    for scene, (obj, env_scene) in SCENES.items():
        tmp = tempfile.mkdtemp()
        repo, scene_dir = make_tree(tmp, scene, obj)
        env = dict(os.environ, STOCS_REPO_PATH=repo, **env_scene)
        subprocess.run([os.path.join(HOST, "model_preprocess"), obj], env=env, capture_output=True, check=True, timeout=600)
        model = read_model(os.path.join(repo, "models", obj, "model_search.ply"))
        # instance mode numbers its bases in an 8-bit segmentation image (src/stocs.cpp:625): at most 255
        instance = os.path.exists(os.path.join(scene_dir, "probability_maps", "edge.png"))
        long_run = run(scene, 1000, 250 if instance else 5000, tmp)
        short = [run(scene, s, 100, tmp) for s in range(1, 21)]
        poses = [np.array(r["pose"]) for r in short if r["pose"] is not None]
        ref = np.array(long_run["pose"]) if long_run["pose"] is not None else None
        for r in short:
            r["adds_to_long_run_m"] = adds(model, np.array(r["pose"]), ref) if (r["pose"] is not None and ref is not None) else None
        # modal cluster of the short runs: the pose with the most other poses within 1 cm ADD-S
        D = np.array([[adds(model, a, b) for b in poses] for a in poses]) if poses else np.zeros((0, 0))
        modal = int(np.argmax((D < 0.01).sum(1))) if len(poses) else -1
        diam = float(np.linalg.norm(model.max(0) - model.min(0)))
        report[scene] = {
            "object": obj, "model_points": int(len(model)), "model_diameter_m": diam,
            "long_run": dict({k: long_run[k] for k in ("bases_valid", "congruent_sets", "transforms", "best_lcp")},
                             bases=250 if instance else 5000),
            "short_runs": len(short), "short_runs_with_pose": len(poses),
            "best_lcp": {"min": min(r["best_lcp"] for r in short), "median": float(np.median([r["best_lcp"] for r in short])),
                         "max": max(r["best_lcp"] for r in short)},
            "transforms": {"min": min(r["transforms"] for r in short), "median": float(np.median([r["transforms"] for r in short])),
                           "max": max(r["transforms"] for r in short)},
            "within_1cm_adds_of_long_run": int(sum(1 for r in short if r["adds_to_long_run_m"] is not None and r["adds_to_long_run_m"] < 0.01)),
            "within_10pct_diameter_adds_of_long_run": int(sum(1 for r in short if r["adds_to_long_run_m"] is not None and r["adds_to_long_run_m"] < 0.1 * diam)),
            "modal_cluster_size_1cm": int((D[modal] < 0.01).sum()) if modal >= 0 else 0,
            "runs": [{k: v for k, v in r.items() if k != "pose"} for r in short]}
        shutil.rmtree(tmp, ignore_errors=True)
        print(scene, json.dumps({k: v for k, v in report[scene].items() if k != "runs"}), flush=True)
    json.dump(report, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
