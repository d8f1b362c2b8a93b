#!/usr/bin/env python
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small committed summaries under profiles/.
usage: python profiles/summarize.py <tag>   (expects gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_k_*.ncu-rep)"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
out = {"tag": tag}

p = os.path.join(G, tag + "_launches.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        if r[vi].strip().lower() in ("nan", "n/a", ""):      # a launch ncu could not time (seen once per run under green contexts)
            continue
        tot[name] += float(r[vi].replace(",", "")) / 1e6
        cnt[name] += 1
    total = sum(tot.values())
    out["launch_list"] = {k: {"launches": cnt[k], "total_ms": round(tot[k], 3), "avg_ms": round(tot[k] / cnt[k], 4),
                              "share": round(tot[k] / total, 4)} for k in sorted(tot, key=tot.get, reverse=True)}
    lines = ["kernel,launches,total_ms,avg_ms,share"] + ["%s,%d,%.3f,%.4f,%.4f" % (k, v["launches"], v["total_ms"], v["avg_ms"], v["share"])
                                                       for k, v in out["launch_list"].items()]
    open(os.path.join(ROOT, "profiles", tag + "_launch_list.csv"), "w").write("\n".join(lines) + "\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
traffic = {}
for k in ("k_model", "k_range", "k_emit", "k_pack", "k_flac", "k_decode", "k_md5", "k_padding"):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, k))
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, v = rows[0], rows[1], rows[-1]
    d = {}
    for i, n in enumerate(h):
        if n in WANT:
            d[n] = "%s %s" % (v[i], units[i])
    out[k] = d
    try:
        def val(name):
            i = h.index(name)
            x = float(v[i].replace(",", ""))
            u = units[i].lower()
            return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        traffic[k] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    except Exception as e:      # noqa: BLE001
        print("traffic", k, e)
json.dump(out, open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.json"), "w"), indent=1)
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:6000])
