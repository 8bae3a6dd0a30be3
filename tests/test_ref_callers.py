"""The reference's two command-line callers, UNMODIFIED, against the drop-in host layer.

CPU part (runs wherever /root/reference exists): src/stocs_match_one_object.cpp and
src/model_preprocess.cpp compile and link against model_matching_b200/host + libstocs_b200.so
(tests/ref_callers/Makefile) -- no source of the reference is copied into this repository.
GPU part: the binaries built here run on the reference's YCB example (the one configuration they
are compiled for, src/stocs_match_one_object.cpp:4-24) and produce the same files and numbers as
this repository's own CLI run with the reference's quad selection (STOCS_REF_SHUFFLE=1)."""
import filecmp
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "model_matching_b200", "host")
BUILD = os.path.join(ROOT, "tests", "ref_callers", "_build")
REF = "/root/reference"
REF_REPO_PATH = "/media/chaitanya/DATADRIVE0/github/model_matching"    # src/stocs_match_one_object.cpp:4


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_unmodified_reference_callers_build_against_the_shim():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "ref_callers"), "clean"], check=True, capture_output=True)
    p = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "ref_callers")], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    for exe in ("stocs_single_ref", "model_preprocess_ref"):
        assert os.access(os.path.join(BUILD, exe), os.X_OK)
    # compiled from the reference tree in place: the command lines name its files, the repo holds no copy
    assert f"{REF}/src/stocs_match_one_object.cpp" in p.stdout and f"{REF}/src/model_preprocess.cpp" in p.stdout
    for dirpath, _, files in os.walk(ROOT):
        if ".git" in dirpath:
            continue
        assert "stocs_match_one_object.cpp" not in files and "model_preprocess.cpp" not in files, dirpath


@pytest.mark.gpu
def test_reference_callers_run_and_agree_with_the_cli(tmp_path):
    for exe in ("stocs_single_ref", "model_preprocess_ref"):
        if not os.path.exists(os.path.join(BUILD, exe)):
            pytest.skip("tests/ref_callers/_build is missing (built by __graft_entry__.build() where /root/reference exists)")
    from test_shim_gpu import make_tree
    out = {}
    for who in ("ref", "own"):
        repo, scene_dir = make_tree(str(tmp_path / who), "ycb", "024_bowl")
        env = dict(os.environ, STOCS_SEED="7")
        if who == "ref":
            env["STOCS_PATH_REMAP"] = f"{REF_REPO_PATH}={repo}"
            pre = [os.path.join(BUILD, "model_preprocess_ref"), "024_bowl"]
            run = [os.path.join(BUILD, "stocs_single_ref"), scene_dir, "024_bowl"]
        else:
            env.update(STOCS_REPO_PATH=repo, STOCS_REF_SHUFFLE="1")
            pre = [os.path.join(HOST, "model_preprocess"), "024_bowl"]
            run = [os.path.join(HOST, "stocs_single"), scene_dir, "024_bowl"]
        p1 = subprocess.run(pre, env=env, capture_output=True, text=True, timeout=300)
        assert p1.returncode == 0, p1.stdout + p1.stderr
        p2 = subprocess.run(run, env=env, capture_output=True, text=True, timeout=300)
        assert p2.returncode == 0, p2.stdout + p2.stderr
        out[who] = (repo, scene_dir, p1.stdout, p2.stdout)
    (rrepo, rscene, rpre, rrun), (orepo, oscene, opre, orun) = out["ref"], out["own"]
    # offline artefacts: byte-identical
    for f in ("model_search.ply", "ppf_map"):
        assert filecmp.cmp(os.path.join(rrepo, "models", "024_bowl", f), os.path.join(orepo, "models", "024_bowl", f), shallow=False), f
    assert "After sampling |M|= 472" in rpre and "After sampling |M|= 472" in opre
    # online: same counters on stdout, same pose file, same debug clouds
    def numbers(text):
        return [re.search(pat, text).group(1) for pat in (r"\|M\| = (\d+)", r"\|map\(M\)\| = (\d+)", r"\|S\|: (\d+)", r"Sampled (\d+) bases",
                                                          r"found (\d+) congruent sets", r"Transforms to verify: (\d+)",
                                                          r"best index: (-?\d+), maximum score: ([0-9.e+-]+)")]
    assert numbers(rrun) == numbers(orun)
    for f in ("best_pose_candidate_024_bowl.txt", "dbg/best_pose.ply", "dbg/scene.ply", "dbg/sampled_scene.ply"):
        assert filecmp.cmp(os.path.join(rscene, f), os.path.join(oscene, f), shallow=False), f
    pose = np.loadtxt(os.path.join(rscene, "best_pose_candidate_024_bowl.txt"))
    assert pose.shape == (12,)
