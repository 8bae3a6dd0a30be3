"""Launches the scoring kernel a few times on one of the two named workloads -- the target of the
ncu captures kept under profiles/ (a number printed under ncu is never a bench value).
    ncu --set full --import-source on --kernel-name regex:score_lcp_kernel --launch-skip 3 --launch-count 1 \
        python profiles/kernel_target.py s1fit 5"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from model_matching_b200 import Context

which, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 5
sc, mpos, mnrm, T = bench.workload(0, bench.H_TOTAL)
ctx = Context(0); ctx.upload_model(mpos, mnrm); ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
if which == "s1fit":
    T, info = bench.fitted_hypotheses(ctx, bench.H_TOTAL)
H = len(T)
dT = torch.from_numpy(T).cuda(); dl = torch.empty(H, dtype=torch.float32, device="cuda"); di = torch.empty(H, dtype=torch.int32, device="cuda")
for _ in range(n):
    ctx.score_lcp_device(dT.data_ptr(), H, dl.data_ptr(), di.data_ptr())
torch.cuda.synchronize()
print(which, H, ctx.kernel_ms_stats())
