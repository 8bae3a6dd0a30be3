"""Cost of the per-step tail (top-K partial + merge/pack [+ probe]) measured alone with CUDA events (development tool)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from model_matching_b200 import Context
sc, mpos, mnrm, T = bench.workload(0, 1000000)
ctx = Context(0); ctx.upload_model(mpos, mnrm); ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])
s = torch.cuda.Stream(); sp = ctypes.c_void_p(s.cuda_stream)
with torch.cuda.stream(s):
    for H in (1000000, 125000):
        dT = torch.from_numpy(T[:H]).cuda(); dl = torch.empty(H, dtype=torch.float32, device="cuda"); di = torch.empty(H, dtype=torch.int32, device="cuda")
        ti = torch.empty(32, dtype=torch.int64, device="cuda"); tv = torch.empty(32, dtype=torch.float32, device="cuda")
        drec = torch.zeros(32 * 64, dtype=torch.uint8, device="cuda")
        ctx.score_lcp_device(dT.data_ptr(), H, dl.data_ptr(), di.data_ptr(), sp)
        def timeit(fn, n=200):
            for _ in range(5): fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(s)
            for _ in range(n): fn()
            b.record(s); torch.cuda.synchronize()
            return a.elapsed_time(b) / n * 1e3
        t_topk = timeit(lambda: ctx.reduce_best_device(dl.data_ptr(), H, 32, 0, ti.data_ptr(), tv.data_ptr(), sp))
        t_step = timeit(lambda: ctx.score_sharded_device(dT.data_ptr(), H, 0, 32, drec.data_ptr(), sp), 50)
        t_score = timeit(lambda: ctx.score_lcp_device(dT.data_ptr(), H, dl.data_ptr(), di.data_ptr(), sp), 50)
        n, kms, _ = ctx.kernel_ms_stats()
        print("H=%d: top-K (2 launches) %.1f us | score call %.1f us (kernel %.1f us) | whole step %.1f us" % (H, t_topk, t_score, kms * 1e3, t_step))
