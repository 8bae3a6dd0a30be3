"""A/B timing of score-kernel build variants on the S1 and S1-fit workloads (development tool).
Build variants with profiles/build_variants.sh, then on the GPU box:
    echo base atom32 w16b3 | python profiles/ab_probe.py
Each variant runs in its own process (STOCS_B200_LIB selects the library); prints ms per 10^6
hypotheses for both workloads and a hash of the results, which must not change between variants.
The S1-fit hypotheses are generated once (first variant) and cached in /tmp."""
import sys, os, time, ctypes, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, '.')
    import numpy as np, torch, bench, hashlib
    from model_matching_b200 import Context
    sc, mpos, mnrm, T = bench.workload(0, 1000000); H = len(T)
    ctx = Context(0); ctx.upload_model(mpos, mnrm); ctx.upload_scene(sc['pos'], sc['nrm'], sc['cls'])
    fitp = '/tmp/s1fit_T.npy'
    if not os.path.exists(fitp):
        Tf, info = bench.fitted_hypotheses(ctx, H)
        np.save(fitp, Tf)
    Tf = np.load(fitp)
    s = torch.cuda.Stream()
    out = '%-16s' % sys.argv[1]
    with torch.cuda.stream(s):
        sp = ctypes.c_void_p(s.cuda_stream)
        for name, TT in (('S1', T), ('S1-fit', Tf), ('S1/8', T[:125000])):
            Hn = len(TT)
            dT = torch.from_numpy(TT).cuda(); dl = torch.empty(Hn, dtype=torch.float32, device='cuda'); di = torch.empty(Hn, dtype=torch.int32, device='cuda')
            for _ in range(3): ctx.score_lcp_device(dT.data_ptr(), Hn, dl.data_ptr(), di.data_ptr(), sp)
            torch.cuda.synchronize(); ctx.kernel_ms_stats(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(20): ctx.score_lcp_device(dT.data_ptr(), Hn, dl.data_ptr(), di.data_ptr(), sp)
            e1.record(s)
            torch.cuda.synchronize(); n, ms, mx = ctx.kernel_ms_stats()
            out += ' | %s kernel %.3f call %.3f ms lcp %s inl %d' % (name, ms, e0.elapsed_time(e1) / 20, hashlib.sha1(dl.cpu().numpy().tobytes()).hexdigest()[:8], int(di.sum()))
    print(out, flush=True)
else:
    for v in sys.stdin.read().split():
        lib, _, scale = v.partition('@')
        lib, _, flag = lib.partition('!')          # name!ENVVAR sets ENVVAR=1 for that run (e.g. lpt!STOCS_NO_LPT)
        env = dict(os.environ, STOCS_B200_LIB=os.path.abspath('gpurun_variants/lib_%s.so' % lib))
        if scale: env['STOCS_CELL_SCALE'] = scale
        if flag: env[flag] = '1'
        subprocess.run([sys.executable, 'profiles/ab_probe.py', v], env=env)
