#!/usr/bin/env python
"""Latency per object pose on the reference's example scenes, alone (bench.py's `pose_latency` extra
without the throughput run), plus -- with --trace -- the per-stage wall clock of one upload and one
pipeline run (STOCS_TRACE=1 synchronises after every stage, so the stages add up to MORE than the
untraced call).   python profiles/pose_latency.py [--trace] > pose_latency.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--trace-child" in sys.argv:
    import numpy as np
    from model_matching_b200 import Context
    with np.load(os.path.join(ROOT, "tests", "golden", "golden_ycb.npz")) as z:
        g = {k: np.ascontiguousarray(z[k]) for k in ("mpos", "mnrm", "spos", "snrm", "scls", "spix")}
    ctx = Context(0)
    for _ in range(3):
        ctx.upload_model(g["mpos"], g["mnrm"]); ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"]); ctx.run_pipeline(1, 100, 200)
    sys.stderr.write("[stocs trace] ---- steady state ----\n")
    ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"])
    ctx.run_pipeline(2, 100, 200)
    sys.exit(0)

import bench
from model_matching_b200 import Context

out = bench.pose_latency(lambda: Context(0), False)
if "--trace" in sys.argv:
    env = dict(os.environ, STOCS_TRACE="1")
    err = subprocess.run([sys.executable, os.path.abspath(__file__), "--trace-child"], env=env, capture_output=True, text=True).stderr
    lines = [l for l in err.splitlines() if l.startswith("[stocs trace]")]
    k = max(i for i, l in enumerate(lines) if "steady state" in l)
    out["stage_trace_ycb"] = [l[len("[stocs trace] "):].strip() for l in lines[k + 1:]]
print(json.dumps(out, indent=1))
