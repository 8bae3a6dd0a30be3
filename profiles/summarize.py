#!/usr/bin/env python
"""Turns the ncu artefacts brought back from the GPU box (gpurun_out/) into the text summaries
committed under profiles/.  Usage: python profiles/summarize.py <prof.ncu-rep> <launches.csv> <bench.json> <tag>"""
import collections
import csv
import json
import shutil
import subprocess
import sys

rep, launches, bench, tag = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def num(s):
    return float(s.replace(",", ""))


def tobytes(v, u):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


lines = ["# ncu --set full --clock-control none --import-source on, score_lcp_kernel, 4th launch of",
         "# `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (S1: 1,048,576 scene points, |M|=512, 1e6 hypotheses), B200"]
for w in want:
    if w in d:
        lines.append(f"{w:70s} {d[w][0]:>22s} {d[w][1]}")
st = [(num(d[h][0]), h) for h in hdr if "stalled" in h and "ratio" in h and "not_issued" not in h and "per_issue_active" in h]
lines.append("# warp stall reasons (warps per issue-active cycle)")
for v, h in sorted(st, reverse=True)[:8]:
    lines.append(f"{h:90s} {v:8.3f}")
rd, wr = tobytes(*d["dram__bytes_read.sum"]), tobytes(*d["dram__bytes_write.sum"])
lines.append(f"# dram traffic per launch = {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g}); "
             "algorithmic bytes per launch = 3.2824e10 (H*(56+64*|M|), SURVEY 8d)")
open(f"profiles/{tag}_score_lcp_kernel_ncu.txt", "w").write("\n".join(lines) + "\n")
json.dump({"kernel": "score_lcp_kernel", "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
           "source": f"profiles/{tag}_score_lcp_kernel_ncu.txt"}, open("profiles/score_kernel_traffic.json", "w"), indent=1)

shutil.copy(launches, f"profiles/{tag}_launches.csv")
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
out = [f"# per-kernel totals from profiles/{tag}_launches.csv (ncu gpu__time_duration.sum, cold-cache, serialised:",
       "# compare SHARES).  Besides the warm-up and timed steps the bench process uploads the scene/model (index build),",
       "# runs the end-to-end steps (zero-copy: 1 launch each; staged: 4 chunked launches each) and the YCB pose-latency",
       "# pipeline, hence the other kernels and the short score_lcp_kernel launches."]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    out.append(f"{k:54s} n={v[0]:4d} total_ms={v[1] / 1e6:10.3f} share={v[1] / tot:.4f}")
# share of the dominant kernel inside ONE bench step (score 10^6 hypotheses + top-32 reduction)
big = [num(r[mv]) for r in rows[hi + 1:] if len(r) > mv and "score_lcp_kernel" in r[kn] and num(r[mv]) > 1.5e6]
tp = [num(r[mv]) for r in rows[hi + 1:] if len(r) > mv and "topk_partial_kernel" in r[kn]]
tm = [num(r[mv]) for r in rows[hi + 1:] if len(r) > mv and "topk_merge_kernel" in r[kn]]
if big and tp and tm:
    mean = lambda v: sum(v) / len(v)
    step = mean(big) + mean(tp) + mean(tm)
    bj = json.load(open(bench))
    out.append("# one bench step = score_lcp_kernel (10^6 hypotheses) + topk_partial + topk_merge:")
    out.append(f"#   ncu: {mean(big) / 1e6:.3f} + {mean(tp) / 1e6:.3f} + {mean(tm) / 1e6:.3f} ms -> score share {mean(big) / step:.3f}")
    out.append(f"#   bench.py (CUDA events): kernel_ms {bj['roofline']['kernel_ms']:.3f} of ms_per_step {bj['ms_per_step']:.3f}"
               f" -> share {bj['roofline']['kernel_ms'] / bj['ms_per_step']:.3f}")
open(f"profiles/{tag}_launch_shares.txt", "w").write("\n".join(out) + "\n")
import os
if os.path.abspath(bench) != os.path.abspath(f"profiles/{tag}_bench_n1.json"):
    shutil.copy(bench, f"profiles/{tag}_bench_n1.json")
print("\n".join(lines[-14:]))
print("\n".join(out[:12]))
