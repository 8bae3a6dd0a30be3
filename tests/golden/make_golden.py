#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ with the CPU oracle -- nothing here needs a GPU.

  python tests/golden/make_golden.py [ycb linemod packed] [--check-dump gpurun_out/r2b/inputs]

* golden_synth.npz   -- outputs of every stage on the seeded synthetic workloads of tests/scenes.py
                        and tests/conftest.py (inputs are regenerated from their seeds).
* golden_<scene>.npz -- for the reference's three example scenes: the INPUTS of the hot path (scene
                        point set = the oracle's restatement of rgbd::load_rgbd_data_sampled on the
                        example frame; model point set = the host C++ restatement of
                        pre_process_model's PCL half, model_matching_b200/host/test_host_model, CPU
                        only) and the oracle's OUTPUT of every stage on them.  `packed` runs the
                        instance-mode sampler (edge map present), the other two the class-mode one.
  --check-dump DIR:     compare those inputs with what stocs_single uploaded on the GPU box
                        (tests/golden/dump_inputs.sh): they must be bit-identical.
The reference ships no expected outputs; these pin the ORACLE so that both it and the CUDA path
are regression-checked against fixed numbers ("parity unpinned" w.r.t. the reference itself).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle  # noqa: E402
from model_matching_b200 import synth  # noqa: E402
from scenes import object_scene  # noqa: E402

SEED = 20181018


def stage_outputs(sc_pos, sc_nrm, sc_cls, mpos, mnrm, n_bases, n_random, max_quads=60, spix=None, edge=None):
    omap = oracle.PPFMap(mpos, mnrm)
    est = oracle.Estimator(sc_pos, sc_nrm, sc_cls, mpos, mnrm, ppfmap=omap, spix=spix)
    out = {"map_keys": np.int64(omap.num_keys), "map_entries": np.int64(omap.num_entries)}
    ok, ids, inv = [], [], []
    if edge is not None:      # instance mode (src/stocs.cpp:559-751): stateful sequence, base numbers 1..n
        import hashlib
        est.set_edge_map(edge)
        mask_sha = []
        for b in range(n_bases):
            o, i, v, _, mask = est.sample_instance_base(SEED, b + 1, 0.9)
            ok.append(o); ids.append(i); inv.append(v)
            mask_sha.append(np.frombuffer(hashlib.sha1(mask.tobytes()).digest()[:8], np.uint64)[0])
        out["mask_sha"] = np.array(mask_sha, np.uint64)
        out["class_prob_after"] = est.class_prob()      # decayed priors: the scores below use them
    else:
        for b in range(n_bases):
            o, i, v, _ = est.sample_class_base(SEED, b)
            ok.append(o); ids.append(i); inv.append(v)
    out["base_ok"], out["base_ids"], out["base_inv"] = np.array(ok), np.array(ids), np.array(inv)
    quads, offs, Ts = [], [0], []
    for b in range(n_bases):
        if not ok[b]:
            offs.append(offs[-1]); continue
        q, nP, nQ = est.find_congruent(ids[b], inv[b][0], inv[b][1])
        quads.append(q); offs.append(offs[-1] + len(q))
        for qq in q[:max_quads]:
            f, Tc, Tw = est.fit(ids[b], qq)
            if f:
                Ts.append(Tc)
    out["quads"] = np.concatenate(quads) if quads else np.zeros((0, 4), np.int32)
    out["quad_offsets"] = np.array(offs, np.int64)
    # hypotheses: the fitted transforms + random ones
    rng = np.random.Generator(np.random.Philox(99))
    s, m = est.centred()
    R = synth.random_rotations(rng, n_random)
    t = rng.uniform(s.min(0), s.max(0), size=(n_random, 3))
    T = np.concatenate([np.array(Ts, np.float32).reshape(-1, 16), synth.to_colmajor16(R, t)])
    lcp, inl = est.score(T, threads=os.cpu_count() or 1)
    out["T"], out["lcp"], out["inliers"] = T, lcp, inl
    bi, bl = oracle.best(lcp)
    out["best_index"], out["best_lcp"] = np.int64(bi), np.float32(bl)
    cs, cm = est.centroids()
    out["centroid_scene"], out["centroid_model"] = cs, cm
    return out


def read_inputs(path):
    raw = open(path, "rb").read()
    S, M = np.frombuffer(raw[:16], np.int64)
    o = 16
    def take(n, dt):
        nonlocal o
        a = np.frombuffer(raw[o:o + n * 4], dt).copy(); o += n * 4
        return a
    d = dict(spos=take(S * 3, np.float32).reshape(S, 3), snrm=take(S * 3, np.float32).reshape(S, 3),
             scls=take(S, np.float32), spix=take(S * 2, np.int32).reshape(S, 2),
             mpos=take(M * 3, np.float32).reshape(M, 3), mnrm=take(M * 3, np.float32).reshape(M, 3))
    assert o == len(raw)
    return d


# (object, intrinsics {fx,cx,fy,cy}, depth scale, model normal radius / read scale / voxel) -- the reference's
# settings for its three examples (src/stocs_match_one_object.cpp:7-24, README.md:47-69)
FRAMES = {
    "ycb": ("024_bowl", [1066.778, 312.986, 1067.487, 241.310], 1 / 10000.0, (0.005, 1.0, 0.01)),
    "linemod": ("obj_06", [572.4114, 325.2611, 573.57043, 242.04899], 1 / 1000.0, (5, 0.001, 10)),
    "packed": ("dove", [615.957763671875, 308.1098937988281, 615.9578247070312, 246.33352661132812], 0.000125, (0.005, 1.0, 0.005)),
}


def frame_inputs(scene):
    """the point sets stocs_single uploads for an example scene, computed on the CPU"""
    import subprocess
    import tempfile
    import cv2
    obj, K, depth_scale, (radius, mscale, mvoxel) = FRAMES[scene]
    d = os.path.join(HERE, "examples", scene)
    depth = cv2.imread(os.path.join(d, "depth.png"), cv2.IMREAD_UNCHANGED)
    bgr = cv2.imread(os.path.join(d, "rgb.png"), cv2.IMREAD_COLOR)
    prob = cv2.imread(os.path.join(d, "probability_maps", obj + ".png"), cv2.IMREAD_UNCHANGED)
    ef = os.path.join(d, "probability_maps", "edge.png")
    edge = cv2.imread(ef, cv2.IMREAD_GRAYSCALE) if os.path.exists(ef) else None
    sc = oracle.build_scene_cloud(depth, bgr, prob, edge, K, np.float32(depth_scale), 0.005, 0.10)
    exe = os.path.join(HERE, "..", "..", "model_matching_b200", "host", "test_host_model")
    out = os.path.join(tempfile.mkdtemp(), "model_search.ply")
    subprocess.run([exe, os.path.join(HERE, "models", obj, "textured_vertices.ply"), str(radius), str(mscale), str(mvoxel), out],
                   check=True, capture_output=True)
    with open(out) as f:
        for line in f:
            if line.startswith("element vertex"):
                n = int(line.split()[2])
            if line.startswith("end_header"):
                break
        a = np.loadtxt(f, max_rows=n, dtype=np.float64)
    mpos, mn = a[:, :3].astype(np.float32), a[:, 6:9].astype(np.float32)
    # Point3D::set_normal re-normalises on load (include/point3d.hpp:43-45), binary32, a + (b + c)
    nn = np.sqrt(mn[:, 0] * mn[:, 0] + (mn[:, 1] * mn[:, 1] + mn[:, 2] * mn[:, 2]))
    mnrm = (mn / nn[:, None]).astype(np.float32)
    return dict(spos=sc["pos"], snrm=sc["nrm"], scls=sc["cls"], spix=sc["pix"], mpos=mpos, mnrm=mnrm), edge


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    check = sys.argv[sys.argv.index("--check-dump") + 1] if "--check-dump" in sys.argv else None
    if check:
        args = [a for a in args if a != check]
    if not args or "synth" in args:
        sc, mpos, mnrm = object_scene()
        g = stage_outputs(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm, n_bases=48, n_random=3000)
        np.savez_compressed(os.path.join(HERE, "golden_synth.npz"), **g)
        print("golden_synth", {k: (v.shape if hasattr(v, "shape") else v) for k, v in g.items()})
    for name in args:
        if name == "synth":
            continue
        d, edge = frame_inputs(name)
        if check:
            dump = read_inputs(os.path.join(check, f"{name}_inputs.bin"))
            for k in d:
                assert np.array_equal(d[k], dump[k]), (name, k)
            print(f"{name}: CPU-derived inputs == the GPU box's upload dump (bit-identical)")
        g = stage_outputs(d["spos"], d["snrm"], d["scls"], d["mpos"], d["mnrm"], n_bases=32, n_random=8000, max_quads=100,
                          spix=d["spix"], edge=edge)
        g.update(d)
        np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **g)
        print(f"golden_{name}", "S", len(d["spos"]), "M", len(d["mpos"]), "keys", g["map_keys"], "entries", g["map_entries"],
              "bases ok", int(g["base_ok"].sum()), "quads", len(g["quads"]), "H", len(g["T"]), "best", g["best_index"], g["best_lcp"],
              "max inl", g["inliers"].max())


if __name__ == "__main__":
    main()
