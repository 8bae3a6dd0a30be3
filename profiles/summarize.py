#!/usr/bin/env python
"""Turns the ncu artefacts brought back from the GPU box (gpurun_out/) into the text summaries
committed under profiles/ (run here, on the CPU box: `ncu -i` only reads reports).

    python profiles/summarize.py <dir with score_s1.ncu-rep, score_s1fit.ncu-rep, launches.csv> <bench.json> <tag>
                                 [--before <old .ncu-rep>]

Writes profiles/<tag>_score_lcp_kernel_ncu_{s1,s1fit}.txt (key counters + stall reasons),
profiles/<tag>_score_lcp_source_lines.txt (shared-memory wavefronts / local-memory requests / instructions
by source line of score.cu, with the --before report beside it), profiles/<tag>_launches.csv + _launch_shares.txt,
and profiles/score_kernel_traffic.json (DRAM bytes per launch, stamped with the sources they were measured on)."""
import collections
import csv
import datetime
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def num(s):
    return float(s.replace(",", ""))


def tobytes(v, u):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    vals = [r for r in rows[2:] if "score_lcp" in " ".join(r[:8])][0]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def source_lines(rep):
    """per CUDA source line of score.cu: [line no, text, warp instructions, shared wavefronts, local sectors, samples]
    (ncu --page source --print-source cuda,sass: the rows that carry a line number hold that line's totals)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Line No"), None)
    if hi is None:
        return None
    h = rows[hi]
    ci = h.index("Instructions Executed")
    csh = h.index("L1 Wavefronts Shared")
    cl = h.index("L2 Theoretical Sectors Local")
    cs = h.index("# Samples")
    res = []
    for r in rows[hi + 1:]:
        if len(r) > max(ci, csh, cl) and r[0].strip().isdigit():
            try:
                res.append([int(r[0]), r[1].strip(), num(r[ci]), num(r[csh]), num(r[cl]), num(r[cs])])
            except ValueError:
                pass
    return res


def summarize_kernel(rep, tag, wl, workload_text):
    d = raw_page(rep)
    lines = [f"# ncu --set full --clock-control none --import-source on -k regex:score_lcp_kernel -s 3 -c 1, B200",
             f"# target: python profiles/kernel_target.py {wl} 5  ({workload_text})"]
    for w in WANT:
        if w in d:
            lines.append(f"{w:92s} {d[w][0]:>22s} {d[w][1]}")
    st = [(num(d[h][0]), h) for h in d if "stalled" in h and "ratio" in h and "not_issued" not in h and "per_issue_active" in h]
    lines.append("# warp stall reasons (warps per issue-active cycle)")
    for v, h in sorted(st, reverse=True)[:8]:
        lines.append(f"{h:92s} {v:8.3f}")
    rd, wr = tobytes(*d["dram__bytes_read.sum"]), tobytes(*d["dram__bytes_write.sum"])
    lines.append(f"# dram traffic per launch = {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g}); "
                 "algorithmic bytes per launch (contract) = 3.2824e10 (H*(56+64*|M|), SURVEY 8d)")
    path = f"profiles/{tag}_score_lcp_kernel_ncu_{wl}.txt"
    open(os.path.join(ROOT, path), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[2:14]))
    return {"kernel": "score_lcp_kernel", "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
            "kernel_ms_under_ncu": num(d["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(d["gpu__time_duration.sum"][1], 1),
            "hypotheses_per_launch": 1000000, "workload": workload_text, "file": path}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    before = sys.argv[sys.argv.index("--before") + 1] if "--before" in sys.argv else None
    if before:
        args = [a for a in args if a != before]
    d, bench, tag = args[:3]
    import bench as benchmod
    stamp = benchmod.source_stamp()
    traffic = {"source_stamp": stamp, "captured": datetime.date.today().isoformat(),
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE score_lcp_kernel launch (ncu --set full); bench.py quotes "
                       "an entry only for the workload it was captured on and flags it stale when csrc sources changed since"}
    for wl, text in (("s1", "S1: 1,048,576-point scene, |M|=512, 1e6 hypotheses, 1% near-truth"),
                     ("s1fit", "S1-fit: same scene and model, 1e6 hypotheses fitted by the pipeline")):
        rep = os.path.join(d, f"score_{wl}.ncu-rep")
        if os.path.exists(rep):
            t = summarize_kernel(rep, tag, wl, text)
            t["source_stamp"] = stamp
            t["captured"] = traffic["captured"]
            traffic["s1" if wl == "s1" else "s1_fit"] = t
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json"), "w"), indent=1)

    # ---- source-line table (instructions, shared wavefronts, local sectors) after [and before]
    try:
        after = source_lines(os.path.join(d, "score_s1.ncu-rep"))
        bef = source_lines(before) if before else None
        if after:
            out = ["# score_lcp_kernel, S1 workload, per SOURCE LINE of csrc/score.cu (ncu --page source --print-source cuda,sass).",
                   "# columns: line | warp instructions executed | L1 shared-memory wavefronts | L2 sectors of LOCAL memory (spills) | stall samples",
                   f"# 'after' = {tag} capture of the committed kernel; 'before' = the round-1 kernel v11 (line numbers of ITS score.cu)."]

            def block(name, t):
                tot = [sum(r[k] for r in t) for k in (2, 3, 4, 5)]
                out.append(f"## {name}: totals: instructions {tot[0]:.4g}, shared wavefronts {tot[1]:.4g}, local sectors {tot[2]:.4g}, samples {tot[3]:.4g}")
                out.append(f"## {name}: top lines by shared-memory wavefronts")
                for r in sorted(t, key=lambda r: -r[3])[:22]:
                    out.append(f"{r[0]:5d} {r[2]:13.0f} {r[3]:13.0f} {r[4]:11.0f} {r[5]:8.0f} | {r[1][:120]}")
                out.append(f"## {name}: lines with local-memory traffic")
                for r in sorted(t, key=lambda r: -r[4])[:10]:
                    if r[4] > 0:
                        out.append(f"{r[0]:5d} {r[2]:13.0f} {r[3]:13.0f} {r[4]:11.0f} {r[5]:8.0f} | {r[1][:120]}")
            block("after", after)
            if bef:
                block("before", bef)
            open(os.path.join(ROOT, f"profiles/{tag}_score_lcp_source_lines.txt"), "w").write("\n".join(out) + "\n")
            print("\n".join(out[:16]))
    except Exception as e:  # the source page layout is version dependent; the raw-counter summaries do not depend on it
        print("source-line table skipped:", repr(e))

    # ---- launch list
    launches = os.path.join(d, "launches.csv")
    shutil.copy(launches, os.path.join(ROOT, f"profiles/{tag}_launches.csv"))
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) > mv:
            name = r[kn].split("(")[0][-52:]
            agg[name][0] += 1
            agg[name][1] += num(r[mv])
    tot = sum(v[1] for v in agg.values())
    out = [f"# per-kernel totals from profiles/{tag}_launches.csv (ncu gpu__time_duration.sum, cold-cache, serialised: compare SHARES)",
           "# command: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras  (scene/model upload, 3 warm-up + 2 timed steps,",
           "# the single-GPU check of the merged records, the end-to-end steps: zero-copy 1 launch each, staged 4 chunked launches each)"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        out.append(f"{k:54s} n={v[0]:4d} total_ms={v[1] / 1e6:10.3f} share={v[1] / tot:.4f}")
    big = [num(r[mv]) for r in rows[hi + 1:] if len(r) > mv and "score_lcp_kernel" in r[kn] and num(r[mv]) > 1.5e6]
    tp = [num(r[mv]) for r in rows[hi + 1:] if len(r) > mv and "topk_kernel" in r[kn]]
    tm = [0.0]
    if big and tp:
        mean = lambda v: sum(v) / len(v)
        step = mean(big) + mean(tp)
        bj = json.load(open(bench))
        out.append("# one bench step at N=1 = score_lcp_kernel (10^6 hypotheses) + topk_kernel (top-32, packs the 64-byte records):")
        out.append(f"#   ncu: {mean(big) / 1e6:.3f} + {mean(tp) / 1e6:.3f} ms -> score share {mean(big) / step:.3f}")
        out.append(f"#   bench.py (CUDA events): kernel_ms {bj['roofline']['kernel_ms']:.3f} of ms_per_step {bj['ms_per_step']:.3f}"
                   f" -> share {bj['roofline']['kernel_ms'] / bj['ms_per_step']:.3f}")
    open(os.path.join(ROOT, f"profiles/{tag}_launch_shares.txt"), "w").write("\n".join(out) + "\n")
    # ---- one frame of the online path (profiles/pose_latency.py --trace-child under the same ncu pass)
    pl = os.path.join(d, "pose_launches.csv")
    if os.path.exists(pl):
        rows = list(csv.reader(open(pl)))
        hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        h = rows[hi]
        kn, mv = h.index("Kernel Name"), h.index("Metric Value")
        data = [(r[kn], num(r[mv])) for r in rows[hi + 1:] if len(r) > mv]
        # the last upload_scene + run_pipeline: from the last pack_attr_kernel on
        start = max(i for i, (k, _) in enumerate(data) if "pack_attr_kernel" in k)
        lines = [f"# kernels of ONE frame on the YCB example (|S| = 13 419, |M| = 472, 100 bases, <= 200 sets per base), steady state:",
                 "# upload_scene (pack_attr .. brick_table) then run_pipeline (sample_bases .. pipe_finalize);",
                 "# ncu gpu__time_duration.sum per launch, cold caches, serialised -- read as a breakdown, not as the call's latency",
                 "# (python profiles/pose_latency.py --trace-child; wall-clock figures: bench.py extra.pose_latency)"]
        tot_u = tot_p = 0.0
        seen_sample = False
        for k, v in data[start:]:
            name = k.split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-60:]
            if "cub::" in k:
                name = "cub::" + k.split("cub::")[1].split("<")[0]
            seen_sample = seen_sample or "sample_bases_kernel" in k
            if seen_sample: tot_p += v
            else: tot_u += v
            lines.append(f"{name:62s} {v / 1e3:9.2f} us")
        lines.append(f"# upload_scene kernels {tot_u / 1e3:.1f} us, run_pipeline kernels {tot_p / 1e3:.1f} us")
        open(os.path.join(ROOT, f"profiles/{tag}_pose_frame_launches.txt"), "w").write("\n".join(lines) + "\n")
    if os.path.abspath(bench) != os.path.abspath(os.path.join(ROOT, f"profiles/{tag}_bench_n1.json")):
        shutil.copy(bench, os.path.join(ROOT, f"profiles/{tag}_bench_n1.json"))
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()
