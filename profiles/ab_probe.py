"""A/B timing of score-kernel build variants on the S1 workload (development tool, not a test).
Build each variant as gpurun_variants/lib_<name>.so (nvcc line of csrc/Makefile plus -D switches,
e.g. -DSCORE_NO_L2_HINTS), then on the GPU box:
    echo base nohints base@0.6 | python profiles/ab_probe.py
Each variant runs in its own process (STOCS_B200_LIB selects the library, name@scale also sets
STOCS_CELL_SCALE); prints ms per 10^6 hypotheses and a hash of the results, which must not change."""
import sys, os, time, ctypes, subprocess, json
if len(sys.argv)>1:
    sys.path.insert(0,'.')
    import numpy as np, torch, bench, hashlib
    from model_matching_b200 import Context
    sc,mpos,mnrm,T=bench.workload(0,1000000); H=len(T)
    ctx=Context(0); ctx.upload_model(mpos,mnrm); ctx.upload_scene(sc['pos'],sc['nrm'],sc['cls'])
    dT=torch.from_numpy(T).cuda(); dl=torch.empty(H,dtype=torch.float32,device='cuda'); di=torch.empty(H,dtype=torch.int32,device='cuda')
    s=torch.cuda.Stream()
    with torch.cuda.stream(s):
        sp=ctypes.c_void_p(s.cuda_stream)
        for _ in range(3): ctx.score_lcp_device(dT.data_ptr(),H,dl.data_ptr(),di.data_ptr(),sp)
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(20): ctx.score_lcp_device(dT.data_ptr(),H,dl.data_ptr(),di.data_ptr(),sp)
        torch.cuda.synchronize(); ms=(time.perf_counter()-t)/20*1e3
    print('%-14s %.3f ms  lcp sha %s inl sum %d'%(sys.argv[1],ms,hashlib.sha1(dl.cpu().numpy().tobytes()).hexdigest()[:10],int(di.sum())),flush=True)
else:
    for v in sys.stdin.read().split():
        lib,_,sc=v.partition('@')
        env=dict(os.environ,STOCS_B200_LIB=os.path.abspath('gpurun_variants/lib_%s.so'%lib))
        if sc: env['STOCS_CELL_SCALE']=sc
        subprocess.run([sys.executable,'profiles/ab_probe.py',v],env=env)
