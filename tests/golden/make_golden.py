#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ with the CPU oracle.

  python tests/golden/make_golden.py [gpurun_out/ycb_inputs.bin gpurun_out/linemod_inputs.bin]

* golden_synth.npz  -- outputs of every stage on the seeded synthetic workloads of tests/scenes.py
                       and tests/conftest.py (inputs are regenerated from their seeds).
* golden_<scene>.npz -- INPUTS (scene / model point sets as stocs_single uploads them for the
                       reference's example scenes, dumped with STOCS_DUMP_INPUTS on the GPU box:
                       the back-projection runs on the GPU) and the oracle's outputs on them.
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


def stage_outputs(sc_pos, sc_nrm, sc_cls, mpos, mnrm, n_bases, n_random, max_quads=60):
    omap = oracle.PPFMap(mpos, mnrm)
    est = oracle.Estimator(sc_pos, sc_nrm, sc_cls, mpos, mnrm, ppfmap=omap)
    out = {"map_keys": np.int64(omap.num_keys), "map_entries": np.int64(omap.num_entries)}
    ok, ids, inv = [], [], []
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


def main():
    sc, mpos, mnrm = object_scene()
    g = stage_outputs(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm, n_bases=48, n_random=3000)
    np.savez_compressed(os.path.join(HERE, "golden_synth.npz"), **g)
    print("golden_synth", {k: (v.shape if hasattr(v, "shape") else v) for k, v in g.items()})
    for path in sys.argv[1:]:
        name = os.path.basename(path).replace("_inputs.bin", "")
        d = read_inputs(path)
        g = stage_outputs(d["spos"], d["snrm"], d["scls"], d["mpos"], d["mnrm"], n_bases=32, n_random=8000, max_quads=100)
        g.update(d)
        np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **g)
        print(f"golden_{name}", "S", len(d["spos"]), "M", len(d["mpos"]), "keys", g["map_keys"], "entries", g["map_entries"],
              "bases ok", int(g["base_ok"].sum()), "quads", len(g["quads"]), "H", len(g["T"]), "best", g["best_index"], g["best_lcp"],
              "max inl", g["inliers"].max())


if __name__ == "__main__":
    main()
