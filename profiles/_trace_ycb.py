import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from model_matching_b200 import Context
z = np.load('tests/golden/golden_ycb.npz'); g = {k: np.ascontiguousarray(z[k]) for k in ("mpos","mnrm","spos","snrm","scls","spix")}
for mode in ("host", "device"):
    if mode == "device": os.environ["STOCS_DEVICE_CENTROID"] = "1"
    ctx = Context(0)
    for _ in range(3):
        ctx.upload_model(g["mpos"], g["mnrm"]); ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"])
    ctx.run_pipeline(1, 100, 200)
    ts = []
    for s in range(2, 12):
        t0 = time.perf_counter(); r = ctx.run_pipeline(s, 100, 200); ts.append(time.perf_counter() - t0)
    print(mode, "pipeline ms", np.round(np.array(ts) * 1e3, 3), "transforms", r.n_transforms, flush=True)
    tu = []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"]); tu.append(time.perf_counter() - t0)
    print(mode, "upload ms", np.round(np.array(tu) * 1e3, 3), flush=True)
    os.environ["STOCS_TRACE"] = "1"
    ctx.run_pipeline(5, 100, 200)
    ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"])
    del os.environ["STOCS_TRACE"]
    ctx.close()
    os.environ.pop("STOCS_DEVICE_CENTROID", None)
