#!/usr/bin/env python
"""profiles/ summaries from an ncu report:  python tools/ncu_summarise.py REPORT.ncu-rep OUT_PREFIX
OUT_PREFIX_kernels.txt : one block per kernel (the launch with the longest duration): time, DRAM bytes read / written and the GB/s they
                         amount to, share of the measured HBM peak, achieved occupancy, issue-slot utilisation, registers, grid, the largest
                         warp-stall reasons per issued instruction
OUT_PREFIX_traffic.json: {kernel: dram bytes per launch} (what bench.py's roofline.traffic reads)"""
import csv
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6554.9
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=0.0):
    try:
        v = float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return default
    u = units[col[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "second": 1, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u)
    return v * scale if scale else v


best = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    t = val(r, "gpu__time_duration.sum")
    if name not in best or t > best[name][0]:
        best[name] = (t, r)
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
lines = [f"# {os.path.basename(rep)}: ncu --set full --clock-control none; per kernel the launch with the longest duration; HBM peak {peak:.1f} GB/s (MEASURED_PEAKS.json)"]
traffic = {}
for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    traffic[name.split("<")[0]] = int(rd + wr)
    st = sorted(((val(r, s), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:4]
    lines.append(f"{name}")
    lines.append(f"    time {t * 1e6:9.1f} us   grid {r[col['launch__grid_size']]} x {r[col['launch__block_size']]}   regs {r[col['launch__registers_per_thread']]}   "
                 f"dram read {rd / 1e6:8.2f} MB  write {wr / 1e6:8.2f} MB  -> {(rd + wr) / max(t, 1e-12) / 1e9:7.1f} GB/s = {(rd + wr) / max(t, 1e-12) / 1e9 / peak:5.3f} of HBM peak")
    lines.append(f"    warps active {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} %   issue slots busy {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} %   "
                 f"warp instructions {val(r, 'smsp__inst_executed.sum'):.3g}   threads/instr {val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):4.1f}   "
                 f"L2 hit {val(r, 'lts__t_sector_hit_rate.pct'):4.1f} %   stalls/issue: " + ", ".join(f"{n} {v:.2f}" for v, n in st))
open(out + "_kernels.txt", "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(out + "_traffic.json", "w"), indent=1, sort_keys=True)
print(f"{len(best)} kernels -> {out}_kernels.txt")
