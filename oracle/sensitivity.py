#!/usr/bin/env python
"""How much do the UNPINNED choices of leaf arithmetic matter?  (TEST INFRASTRUCTURE.)

The reference's results depend on arithmetic that lives in libraries whose sources and versions are
not in its tree (Eigen's evaluation order, libm, overload resolution of unqualified atan2/acos;
SURVEY.md section 8c).  The oracle pins one model (csrc/stocs_math.h).  This script rebuilds the
oracle under each alternative model (the STOCS_MODEL_* switches of stocs_math.h), re-runs every stage
on the committed golden inputs (tests/golden/golden_{synth,ycb,linemod,packed}.npz) and counts the
INTEGER decisions that change: PPF bins of the model table, table keys, sampled bases, congruent
quads, per-hypothesis inlier counts, LCP bit patterns, the winning hypothesis.

    python oracle/sensitivity.py            # builds variants into oracle/_variants/, prints a table,
                                            # writes profiles/r02_arithmetic_sensitivity.json
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VARIANTS = {
    "sum3_left": ["-DSTOCS_MODEL_SUM3_LEFT"],
    "trig_double": ["-DSTOCS_MODEL_TRIG_DOUBLE"],
    "trig_ulp_up": ["-DSTOCS_MODEL_TRIG_ULP=1"],
    "trig_ulp_down": ["-DSTOCS_MODEL_TRIG_ULP=-1"],
    "sum3_left+trig_double": ["-DSTOCS_MODEL_SUM3_LEFT", "-DSTOCS_MODEL_TRIG_DOUBLE"],
}
SEED = 20181018
GOLDENS = ("synth", "ycb", "linemod", "packed")


def build_variant(name, defs):
    out = os.path.join(HERE, "_variants", f"liboracle_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread", "-w", *defs, "-shared", "-o", out,
                           os.path.join(HERE, "stocs_oracle.cpp")])
    return out


def load_inputs(name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    path = os.path.join(ROOT, "tests", "golden", f"golden_{name}.npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as z:
        g = {k: np.ascontiguousarray(z[k]) for k in z.files}
    if name == "synth":
        from scenes import object_scene
        sc, mpos, mnrm = object_scene()
        g.update(spos=sc["pos"], snrm=sc["nrm"], scls=sc["cls"], mpos=mpos, mnrm=mnrm)
    return g


def worker(name):
    """every stage under the library STOCS_ORACLE_LIB selects; prints one JSON object"""
    g = load_inputs(name)
    import oracle
    M = len(g["mpos"])
    i, j = np.meshgrid(np.arange(M), np.arange(M), indexing="ij")
    keep = (i != j).ravel()
    i, j = i.ravel()[keep], j.ravel()[keep]
    ppf = oracle.ppf_compute(g["mpos"][i], g["mnrm"][i], g["mpos"][j], g["mnrm"][j])
    omap = oracle.PPFMap(g["mpos"], g["mnrm"])
    est = oracle.Estimator(g["spos"], g["snrm"], g["scls"], g["mpos"], g["mnrm"], ppfmap=omap,
                           spix=g["spix"] if "spix" in g else None)
    nb = len(g["base_ok"])
    # scene-side PPFs: first point of every golden base against every scene point (the input of the
    # first sampling predicate, src/stocs.cpp:395-407)
    firsts = [int(g["base_ids"][b][0]) for b in range(nb) if g["base_ok"][b]][:8]
    S = len(g["spos"])
    sppf = np.concatenate([oracle.ppf_compute(np.repeat(g["spos"][f][None], S, 0), np.repeat(g["snrm"][f][None], S, 0),
                                               g["spos"], g["snrm"]) for f in firsts])
    if "mask_sha" in g:      # instance mode (packed): the stateful sequence
        import cv2
        est.set_edge_map(cv2.imread(os.path.join(ROOT, "tests", "golden", "examples", name, "probability_maps", "edge.png"),
                                    cv2.IMREAD_GRAYSCALE))
        bases = [est.sample_instance_base(SEED, b + 1, 0.9)[:4] for b in range(nb)]
        est = oracle.Estimator(g["spos"], g["snrm"], g["scls"], g["mpos"], g["mnrm"], ppfmap=omap, spix=g["spix"])
    else:
        bases = [est.sample_class_base(SEED, b) for b in range(nb)]
    # congruent sets on the GOLDEN bases, so that a change in sampling does not hide behind them
    quads = []
    for b in range(nb):
        if g["base_ok"][b]:
            q, _, _ = est.find_congruent(g["base_ids"][b], g["base_inv"][b][0], g["base_inv"][b][1])
            quads.append(q.tolist())
    lcp, inl = est.score(g["T"], threads=os.cpu_count() or 1)
    bi, bl = oracle.best(lcp)
    print(json.dumps({"ppf": ppf.tolist(), "scene_ppf": sppf.tolist(), "map_keys": int(omap.num_keys), "map_entries": int(omap.num_entries),
                      "base_ok": [bool(b[0]) for b in bases], "base_ids": [b[1].tolist() for b in bases],
                      "quads": quads, "inl": inl.tolist(), "lcp": lcp.view(np.uint32).tolist(), "best": int(bi)}))


def run(name, libpath):
    env = dict(os.environ)
    if libpath:
        env["STOCS_ORACLE_LIB"] = libpath
    else:
        env.pop("STOCS_ORACLE_LIB", None)
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], env=env, capture_output=True, text=True, check=True)
    return json.loads(p.stdout.strip().splitlines()[-1])


def compare(a, b):
    pa, pb = np.array(a["ppf"]), np.array(b["ppf"])
    qa = [set(map(tuple, q)) for q in a["quads"]]
    qb = [set(map(tuple, q)) for q in b["quads"]]
    ia, ib = np.array(a["inl"]), np.array(b["inl"])
    okb = [x and y and i == j for x, y, i, j in zip(a["base_ok"], b["base_ok"], a["base_ids"], b["base_ids"])]
    sa, sb = np.array(a["scene_ppf"]), np.array(b["scene_ppf"])
    return {"ppf_pairs": int(len(pa)), "ppf_pairs_changed": int((pa != pb).any(1).sum()),
            "scene_ppf_pairs": int(len(sa)), "scene_ppf_pairs_changed": int((sa != sb).any(1).sum()),
            "map_keys": [a["map_keys"], b["map_keys"]], "map_entries": [a["map_entries"], b["map_entries"]],
            "bases": len(a["base_ok"]), "bases_changed": int(sum(1 for x, y, k in zip(a["base_ok"], b["base_ok"], okb) if (x or y) and not k)),
            "quad_lists": len(qa), "quad_lists_changed": int(sum(x != y for x, y in zip(qa, qb))),
            "quads": int(sum(len(x) for x in qa)), "quads_added_or_removed": int(sum(len(x ^ y) for x, y in zip(qa, qb))),
            "hypotheses": int(len(ia)), "inlier_counts_changed": int((ia != ib).sum()), "max_inlier_delta": int(np.abs(ia - ib).max()),
            "total_inliers": [int(ia.sum()), int(ib.sum())],
            "lcp_bits_changed": int((np.array(a["lcp"]) != np.array(b["lcp"])).sum()),
            "winner": [a["best"], b["best"]], "winner_changed": a["best"] != b["best"]}


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        return worker(sys.argv[2])
    libs = {name: build_variant(name, defs) for name, defs in VARIANTS.items()}
    report = {}
    for gname in GOLDENS:
        if not os.path.exists(os.path.join(ROOT, "tests", "golden", f"golden_{gname}.npz")):
            continue
        base = run(gname, None)
        report[gname] = {v: compare(base, run(gname, lib)) for v, lib in libs.items()}
    out = os.path.join(ROOT, "profiles", "r02_arithmetic_sensitivity.json")
    json.dump(report, open(out, "w"), indent=1)
    cols = ("ppf_pairs_changed", "scene_ppf_pairs_changed", "bases_changed", "quad_lists_changed", "quads_added_or_removed", "inlier_counts_changed",
            "max_inlier_delta", "lcp_bits_changed", "winner_changed")
    print("%-8s %-22s " % ("golden", "model") + " ".join("%22s" % c for c in cols))
    for gname, rows in report.items():
        for v, r in rows.items():
            print("%-8s %-22s " % (gname, v) + " ".join("%22s" % (("%d / %d" % (r[c], r[{"ppf_pairs_changed": "ppf_pairs", "scene_ppf_pairs_changed": "scene_ppf_pairs", "bases_changed": "bases",
                  "quad_lists_changed": "quad_lists", "quads_added_or_removed": "quads", "inlier_counts_changed": "hypotheses",
                  "lcp_bits_changed": "hypotheses"}[c]])) if c in ("ppf_pairs_changed", "scene_ppf_pairs_changed", "bases_changed", "quad_lists_changed",
                  "quads_added_or_removed", "inlier_counts_changed", "lcp_bits_changed") else str(r[c])) for c in cols))
    print("written:", out)


if __name__ == "__main__":
    main()
