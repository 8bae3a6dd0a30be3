#!/usr/bin/env python
"""bench.py -- hypotheses scored / second on the synthetic S1 workload (BASELINE.json configs[3]).

One "step" = one pass of the hot path over one batch: score H = 10^6 rigid-transform hypotheses
(|M| = 512 model points each) against the 1,048,576-point scene, then the best-pose / top-K
reduction (and, for N > 1, the single NCCL all-gather of the per-rank top-K records).
Weak scaling: every rank holds a replica of the scene index and scores its own 10^6 hypotheses.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is produced.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SCENE = 1 << 20
N_MODEL = 512
H_PER_GPU = 1_000_000
TOPK = 32
METRIC = "hypotheses_scored_per_second"
UNIT = "hypotheses/s"


def algorithmic_bytes_per_hypothesis(M):
    # 64 B per NN query (8 cell descriptors x 4 B + one 16 B candidate
    # position + one 16 B attribute record) x |M| queries + 48 B transform in + 8 B results out.
    return 56 + 64 * M


def workload(rank, H):
    from model_matching_b200 import synth
    sc = synth.make_scene(n_points=N_SCENE, seed=1234)
    mpos, mnrm = synth.make_model(N_MODEL)
    T, _ = synth.make_hypotheses(H, sc["pos"], mpos, sc["gt_R"], sc["gt_t"], seed=4321 + rank)
    return sc, mpos, mnrm, T


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started before the warm-up (its
    first sample can take a second); only samples stamped inside the timed region are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.gpu), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_first_sample(self, timeout=3.0):
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = None
                try:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    pass
                rows.append((ts, float(c[1]), float(c[2]),
                             [nm for nm, v in zip(names, c[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t_begin is not None and r[0] is not None and t_begin - 0.05 <= r[0] <= t_end + 0.05]
        window = "timed region"
        if not inside:  # clock skew / unparsable stamps: fall back to every sample (warm-up included)
            inside, window = rows, "whole run"
        sm, smax = [r[1] for r in inside], [r[2] for r in inside]
        reasons = sorted({nm for r in inside for nm in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None, "reasons": reasons,
                "samples": len(sm), "window": window}


def cpu_baseline(sc, mpos, mnrm, T, seconds_target=15.0):
    """The CPU oracle (port of the reference's kd-tree LCP loop) on all host cores, bounded sample."""
    import oracle
    cores = os.cpu_count() or 1
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    n0 = min(len(T), 2000 * cores)
    t0 = time.perf_counter()
    est.score(T[:n0], threads=cores)
    dt = time.perf_counter() - t0
    n = int(min(len(T), max(n0, n0 * seconds_target / max(dt, 1e-6))))
    t0 = time.perf_counter()
    lcp, inl = est.score(T[:n], threads=cores)
    dt = time.perf_counter() - t0
    n1 = min(len(T), 20000)   # the reference itself is single-threaded: report that figure too
    t0 = time.perf_counter()
    est.score(T[:n1], threads=1)
    dt1 = time.perf_counter() - t0
    return est, {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": f"first {n} of the {len(T)} hypotheses of the same workload, {cores} threads, {dt:.1f} s",
                 "single_thread": {"value": n1 / dt1, "unit": UNIT, "sample": f"first {n1} hypotheses, 1 thread, {dt1:.1f} s"}}, (lcp, inl, n)


def pose_latency_fixture(ctx, name, label):
    """GPU-only figures for another of the reference's example scenes (tests/golden/golden_<name>.npz):
    upload + index, model table, fused pipeline, and scoring-only rate on the fixture's own
    hypothesis list (a few thousand transforms: the launch-bound regime of the reference CLI)."""
    path = os.path.join(ROOT, "tests", "golden", f"golden_{name}.npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as z:
        g = {k: np.ascontiguousarray(z[k]) for k in ("mpos", "mnrm", "spos", "snrm", "scls", "spix", "T", "inliers")}
    ctx.upload_model(g["mpos"], g["mnrm"])
    ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"])
    tm, tu = [], []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.upload_model(g["mpos"], g["mnrm"]); tm.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"]); tu.append(time.perf_counter() - t0)
    ctx.run_pipeline(1, 100, 200)
    tp, res = [], None
    for seed in range(2, 12):
        t0 = time.perf_counter(); res = ctx.run_pipeline(seed, 100, 200); tp.append(time.perf_counter() - t0)
    T = np.ascontiguousarray(g["T"], np.float32)
    lcp, inl = ctx.score_lcp(T)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ctx.score_lcp(T); ts.append(time.perf_counter() - t0)
    return {"workload": "%s (|S|=%d, |M|=%d), 100 bases, <=200 sets/base" % (label, len(g["spos"]), len(g["mpos"])),
            "gpu_ms_per_pose": 1e3 * float(np.median(tp)), "gpu_upload_index_ms": 1e3 * float(np.median(tu)),
            "gpu_model_table_ms": 1e3 * float(np.median(tm)), "transforms_scored": int(res.n_transforms),
            "scoring_only_hyp_per_s": len(T) / float(np.median(ts)), "scoring_only_hypotheses": int(len(T)),
            "scoring_matches_fixture": bool(np.array_equal(inl, g["inliers"]))}


def pose_latency(ctx_factory, with_cpu):
    """Secondary BASELINE metric: end-to-end ms per object pose on the reference's YCB example
    (configs[0]): 100 bases -> congruent sets -> <=200 transforms per base -> score -> best, all on
    the device (stocs_b200_run_pipeline), inputs = the scene/model point sets stocs_single uploads
    (tests/golden/golden_ycb.npz).  CPU figure: the oracle on a bounded number of bases, scaled."""
    path = os.path.join(ROOT, "tests", "golden", "golden_ycb.npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as z:  # materialise: an NpzFile decompresses on every access
        g = {k: np.ascontiguousarray(z[k]) for k in ("mpos", "mnrm", "spos", "snrm", "scls", "spix")}
    ctx = ctx_factory()
    ctx.upload_model(g["mpos"], g["mnrm"])                       # first call allocates
    ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"])
    tm, tu = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.upload_model(g["mpos"], g["mnrm"])
        tm.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.upload_scene(g["spos"], g["snrm"], g["scls"], g["spix"])
        tu.append(time.perf_counter() - t0)
    t_model, t_upload = float(np.median(tm)), float(np.median(tu))
    ctx.run_pipeline(1, 100, 200)
    times, res = [], None
    for seed in range(2, 12):
        t0 = time.perf_counter()
        res = ctx.run_pipeline(seed, 100, 200)
        times.append(time.perf_counter() - t0)
    out = {"workload": "YCB 024_bowl example scene (|S|=%d, |M|=%d), 100 bases, <=200 sets/base" % (len(g["spos"]), len(g["mpos"])),
           "gpu_ms_per_pose": 1e3 * float(np.median(times)), "gpu_upload_index_ms": 1e3 * t_upload, "gpu_model_table_ms": 1e3 * t_model,
           "transforms_scored": int(res.n_transforms), "congruent_sets": int(res.n_congruent_sets)}
    try:  # frame -> scene cloud (src/rgbd.cpp:190-279) on the device, PNG decoding excluded
        import cv2
        d = os.path.join(ROOT, "tests", "golden", "examples", "ycb")
        depth = cv2.imread(os.path.join(d, "depth.png"), cv2.IMREAD_UNCHANGED)
        bgr = cv2.imread(os.path.join(d, "rgb.png"), cv2.IMREAD_COLOR)
        prob = cv2.imread(os.path.join(d, "probability_maps", "024_bowl.png"), cv2.IMREAD_UNCHANGED)
        K = [1066.778, 312.986, 1067.487, 241.310]
        ctx.build_scene_cloud(depth, bgr, prob, None, K, 1 / 10000.0, 0.005, 0.10)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            sc = ctx.build_scene_cloud(depth, bgr, prob, None, K, 1 / 10000.0, 0.005, 0.10)
            ts.append(time.perf_counter() - t0)
        out["gpu_scene_cloud_ms"] = 1e3 * float(np.median(ts))
        out["scene_cloud_points"] = int(len(sc["pos"]))
    except Exception as e:  # cv2 missing or data absent: the headline numbers do not depend on it
        out["gpu_scene_cloud_ms"] = None
        out["scene_cloud_error"] = str(e)[:100]
    try:  # a1 alone: synthetic 640x480 depth frame, plane at 1 m + 3 boxes + 1 mm noise
        rng = np.random.Generator(np.random.Philox(1234))
        z = np.full((480, 640), 1.0)
        for (r0, r1, c0, c1, zz) in ((100, 220, 80, 240, 0.80), (250, 400, 300, 420, 0.70), (60, 160, 450, 600, 0.85)):
            z[r0:r1, c0:c1] = zz
        depth16 = np.clip((z + rng.normal(0, 0.001, z.shape)) * 1000.0, 0, 65535).astype(np.uint16)
        bgr8 = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        kk = (572.4114, 325.2611, 573.57043, 242.04899)
        xyz, _ = ctx.backproject(depth16, bgr8, *kk, 0.001)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); ctx.backproject(depth16, bgr8, *kk, 0.001); ts.append(time.perf_counter() - t0)
        tb = float(np.median(ts))
        out["backproject"] = {"frame": "640x480 synthetic (Philox 1234), host buffers in and out", "ms": 1e3 * tb,
                              "pixels_per_s": 307200 / tb, "algorithmic_GBps": 21 * 307200 / tb / 1e9,
                              "note": "21 B per pixel (SURVEY 8d); the call is PCIe/launch bound: 1.5 MB in, 4.9 MB out"}
        if with_cpu:
            import oracle
            oxyz, _ = oracle.backproject(depth16, bgr8, *kk, 0.001)
            out["backproject"]["bit_exact_vs_oracle"] = bool(np.array_equal(xyz.view(np.uint32), np.asarray(oxyz, np.float32).reshape(xyz.shape).view(np.uint32)))
    except Exception as e:
        out["backproject_error"] = str(e)[:100]
    try:  # the other class-mode example of the reference (configs[1] of its README: LINEMOD obj_06)
        out["linemod"] = pose_latency_fixture(ctx, "linemod", "LINEMOD obj_06 example scene")
        out["ycb_scoring_only"] = {k: v for k, v in (pose_latency_fixture(ctx, "ycb", "YCB 024_bowl example scene") or {}).items()
                                   if k.startswith("scoring")}
    except Exception as e:
        out["linemod_error"] = str(e)[:100]
    ctx.close()
    if with_cpu:
        import oracle
        omap = oracle.PPFMap(g["mpos"], g["mnrm"])
        est = oracle.Estimator(g["spos"], g["snrm"], g["scls"], g["mpos"], g["mnrm"], ppfmap=omap)
        nb = 8
        t0 = time.perf_counter()
        Ts = []
        for b in range(nb):
            ok, ids, inv, _ = est.sample_class_base(2, b)
            if not ok:
                continue
            q, _, _ = est.find_congruent(ids, inv[0], inv[1])
            sel = range(len(q)) if len(q) < 200 else [(k * len(q)) // 200 for k in range(200)]
            Ts += [est.fit(ids, q[k])[1] for k in sel]
        if Ts:
            est.score(np.array(Ts, np.float32), threads=1)
        dt = time.perf_counter() - t0
        out["cpu_ms_per_pose"] = 1e3 * dt * 100 / nb
        out["cpu_sample"] = f"oracle, 1 thread, {nb} of 100 bases timed and scaled (PPF map build excluded)"
    return out


def bind_to_gpu_numa_node(local):
    """Multi-rank runs: pin this process to the CPUs next to its GPU (sysfs local_cpulist of the
    PCI device) so that the pinned host buffers the kernels read in place are allocated on that
    NUMA node.  Best effort: returns the CPU list used, or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        devid = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, devid)
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the reference
    itself cannot be compiled here: no Eigen/PCL/OpenCV/Boost) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    sc, mpos, mnrm, T = workload(0, 200_000)
    est = oracle.Estimator(sc["pos"], sc["nrm"], sc["cls"], mpos, mnrm)
    n = min(2500 * cores, len(T) // 2)  # per step; sized so that steps+warmup end within a few minutes
    for w in range(args.warmup):
        est.score(T[:n // 4], threads=cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        off = (k * n) % (len(T) - n)
        est.score(T[off:off + n], threads=cores)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "S1: 1,048,576-point synthetic scene, |M|=512, 1e6 hypotheses/GPU (1% near-truth)",
                      "scene_points": N_SCENE, "model_points": N_MODEL, "hypotheses_per_gpu": H_PER_GPU},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} hypotheses per step ({args.steps} steps) of the S1 workload, {cores} threads"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hyp", type=int, default=H_PER_GPU, help="hypotheses per GPU (default: the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from model_matching_b200 import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libstocs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version / debug lines to stdout by default; stdout carries the JSON line.
        # (NCCL_DEBUG_FILE is honoured only above level VERSION, so VERSION is raised to WARN.)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    H = args.hyp
    warm = max(args.warmup, 3)

    sc, mpos, mnrm, T = workload(rank, H)
    ctx = Context(local)
    ctx.upload_model(mpos, mnrm)
    ctx.upload_scene(sc["pos"], sc["nrm"], sc["cls"])

    # a dedicated (non-default) stream: the C ABI treats a NULL stream as "the context's own"
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    dT = torch.from_numpy(T).to(dev)
    dlcp = torch.empty(H, dtype=torch.float32, device=dev)
    dinl = torch.empty(H, dtype=torch.int32, device=dev)
    dtop_i = torch.empty(TOPK, dtype=torch.int64, device=dev)
    dtop_v = torch.empty(TOPK, dtype=torch.float32, device=dev)
    gather_i = torch.empty(world * TOPK, dtype=torch.int64, device=dev) if world > 1 else None
    gather_v = torch.empty(world * TOPK, dtype=torch.float32, device=dev) if world > 1 else None

    def step(ev=None):
        if ev is not None:
            ev[0].record(stream)
        ctx.score_lcp_device(dT.data_ptr(), H, dlcp.data_ptr(), dinl.data_ptr(), sptr)
        if ev is not None:
            ev[1].record(stream)
        ctx.reduce_best_device(dlcp.data_ptr(), H, TOPK, rank * H, dtop_i.data_ptr(), dtop_v.data_ptr(), sptr)
        if world > 1:  # the path's one collective: all-gather of K (index, lcp) records per rank
            dist.all_gather_into_tensor(gather_i, dtop_i)
            dist.all_gather_into_tensor(gather_v, dtop_v)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(warm):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)

    t_begin = time.time()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        step(kev[k])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop(t_begin, time.time()) if sampler else None
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms_max = float(t[0]), float(t[1])
    value = world * H * args.steps / (ms_total * 1e-3)

    # ---- e2e: the drop-in host-buffer call (pinned host memory in, results out), every step
    hT = torch.from_numpy(T).pin_memory()
    hlcp = torch.empty(H, dtype=torch.float32).pin_memory()
    hinl = torch.empty(H, dtype=torch.int32).pin_memory()
    e2e_steps = max(1, min(args.steps, 20))

    def e2e_step():
        ctx.score_lcp_ptr(hT.data_ptr(), H, hlcp.data_ptr(), hinl.data_ptr())
        return ctx.reduce_best(None, K=TOPK)

    def e2e_measure():
        e2e_step()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            b = e2e_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * H * e2e_steps / float(te[0]), b

    # headline: the kernel reads the page-locked transforms in place over PCIe (64 B per hypothesis
    # cross the bus inside the timed region, no staging copy); second figure: the same call with
    # the transforms staged through HBM by chunked cudaMemcpyAsync (what pageable callers get)
    e2e_value, best = e2e_measure()
    os.environ["STOCS_NO_ZERO_COPY"] = "1"
    e2e_staged, best_staged = e2e_measure()
    del os.environ["STOCS_NO_ZERO_COPY"]
    assert best_staged[0] == best[0] and best_staged[1] == best[1]

    if rank == 0:
        M = N_MODEL
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg_bytes = H * algorithmic_bytes_per_hypothesis(M)
        achieved = alg_bytes / (kernel_ms_max * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "score_kernel_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "S1: 1,048,576-point synthetic scene, |M|=512, 1e6 hypotheses/GPU (1% near-truth)",
                          "scene_points": N_SCENE, "model_points": M, "hypotheses_per_gpu": H,
                          "parallelism": f"hypothesis-sharded x{world}, replicated scene index, one all-gather of top-{TOPK}",
                          "l2": "inputs larger than L2 (64 MB transforms + ~%d MB scene index per step)" % int(
                              (ctx.counters()[3] * 16 + ctx.counters()[2] * 4 + N_SCENE * 16) / 1e6)},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": H * 64, "d2h_bytes_per_step": H * 8,
                       "steps": e2e_steps, "best_index": int(best[0]), "best_lcp": float(best[1]),
                       "input_path": "pinned host transforms read in place by the kernel over PCIe (zero-copy), "
                                     "results copied back with cudaMemcpyAsync",
                       "staged_copy_value": e2e_staged, "numa_cpus": numa},
               "gpu_launches": args.steps * 3,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                            "kernel": "score_lcp_kernel", "kernel_ms": kernel_ms_max,
                            "algorithmic_bytes_per_launch": alg_bytes},
               "clocks": clocks,
               "ties_resolved_by_kdtree": int(ctx.counters()[1])}
        out["roofline"]["note"] = ("algorithmic bytes follow SURVEY 8(d): 64 B for EVERY (hypothesis, model point) query; "
                                   "the kernel answers most queries from a shared-memory occupancy bitmap and a 16 B brick "
                                   "record, so achieved/peak above 1 is expected; `traffic` is the measured DRAM bytes per launch "
                                   "(ncu), `dram_frac` = traffic / kernel time / peak; the ncu capture puts the L1 data pipe "
                                   "at 91 % and instruction issue at 80 %: that is the binding limit")
        if traffic:
            out["roofline"]["dram_achieved"] = traffic / (kernel_ms_max * 1e-3) / 1e9
            out["roofline"]["dram_frac"] = out["roofline"]["dram_achieved"] / peak
        if world == 1:
            out["pose_latency"] = pose_latency(lambda: Context(local), not args.no_cpu_baseline)
        if world == 1 and not args.no_cpu_baseline:
            est, cb, (olcp, oinl, n) = cpu_baseline(sc, mpos, mnrm, T)
            out["cpu_baseline"] = cb
            # parity gate run with the measurement: the sample must match bit for bit
            glcp, ginl = hlcp.numpy()[:n], hinl.numpy()[:n]
            out["parity"] = {"checked": int(n), "inliers_equal": bool(np.array_equal(ginl, oinl)),
                             "lcp_bits_equal": bool(np.array_equal(glcp.view(np.uint32), olcp.view(np.uint32)))}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
