#!/usr/bin/env python
"""Turns gpurun_out/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum) and
gpurun_out/<tag>_prof.ncu-rep (ncu --set full) into profiles/<tag>_summary.md and updates
profiles/traffic.json (DRAM bytes per launch of each of our kernels, read by bench.py).

    python profiles/summarise.py <tag> ["free-text note"]
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
out = [f"# ncu summary `{tag}`", "", note, ""]

lp = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(l for l in open(lp) if l.startswith('"')))
    h = rows[0]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        agg.setdefault(name, {"t": [], "grid": r[gi], "block": r[bi]})["t"].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v["t"]) for v in agg.values())
    out += ["## Launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: "
            "compare shares)", "", "| kernel | launches | avg us | share of GPU time | grid | block |", "|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["t"])):
        out.append(f"| `{k}` | {len(v['t'])} | {sum(v['t']) / len(v['t']) / 1e3:.2f} | {sum(v['t']) / tot:.3f} | {v['grid']} | {v['block']} |")
    out.append("")

rp = os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")
traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
            "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
    idx = [(w, h.index(w)) for w in want if w in h]
    kn = h.index("Kernel Name")
    out += ["## `--set full` capture (one launch each, ~40 replays: durations are not bench values)", ""]
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "")
        out.append(f"### `{name}`")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        vals = {}
        for w, i in idx:
            out.append(f"| {w} | {r[i]} | {units[i]} |")
            vals[w] = (r[i], units[i])
        def to_bytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        try:
            tb = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
            key = name.split("::")[-1].split("<")[0]
            traffic[key] = tb
            out.append(f"| **dram read+write per launch** | {tb / 1e6:.2f} | MB |")
        except Exception:
            pass
        out.append("")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    pipes = [("ALU %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
             ("FMA %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
             ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
             ("LSU %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
             ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             ("warp instructions", "smsp__inst_executed.sum"),
             ("long-scoreboard stall (warps per issue)", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")]
    pipes = [(n, h.index(m)) for n, m in pipes if m in h]
    if pipes:
        out += ["## Pipe utilisation (same capture)", "", "| kernel | " + " | ".join(n for n, _ in pipes) + " |",
                "|---|" + "---|" * len(pipes)]
        for r in rows[2:]:
            name = r[kn].split("(")[0].replace("void ", "")
            out.append(f"| `{name}` | " + " | ".join(r[i] for _, i in pipes) + " |")
        out.append("")

open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
