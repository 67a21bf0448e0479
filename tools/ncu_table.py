#!/usr/bin/env python
"""Summarise an ncu report per kernel: duration, DRAM bytes, achieved GB/s, occupancy, registers.

  ncu -i gpurun_out/r02_stages_batch.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_table.py /tmp/raw.csv [peak_gbs]
"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6551.7
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "occ": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "grid": "launch__grid_size", "block": "launch__block_size", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1_sect_per_req": "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio"}
units = rows[1]


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def scale(v, u):  # to bytes / microseconds
    if v is None:
        return None
    u = u.lower()
    for k, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e3), ("usecond", 1.0), ("nsecond", 1e-3), ("second", 1e6)):
        if u.startswith(k):
            return v * f
    return v


agg = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    m = re.search(r"stage_kernel(?:_occ)?<b2g::(\w+)>", name)
    short = m.group(1) if m else re.sub(r"\(.*", "", name).split("::")[-1][:48]
    d = {}
    for k, metric in want.items():
        if metric in col:
            d[k] = scale(num(r[col[metric]]), units[col[metric]])
    a = agg.setdefault(short, {"n": 0, "dur": 0.0, "rd": 0.0, "wr": 0.0, "occ": [], "regs": d.get("regs"), "grid": d.get("grid"),
                               "block": d.get("block"), "dram_pct": [], "sm_pct": [], "l1": []})
    a["n"] += 1
    a["dur"] += d.get("dur") or 0.0
    a["rd"] += d.get("rd") or 0.0
    a["wr"] += d.get("wr") or 0.0
    for k, dst in (("occ", "occ"), ("dram_pct", "dram_pct"), ("sm_pct", "sm_pct"), ("l1_sect_per_req", "l1")):
        if d.get(k) is not None:
            a[dst].append(d[k])
    a["grid"] = max(a["grid"] or 0, d.get("grid") or 0)
print("| kernel | launches | us / launch | DRAM MB / launch (rd + wr) | GB/s | of %.0f | warps active %% | dram %% | sm %% | regs | grid x block | L1 sectors/req |" % peak)
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
out = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["dur"]):
    n = a["n"]
    us = a["dur"] / n
    mb = (a["rd"] + a["wr"]) / n / 1e6
    gbs = (a["rd"] + a["wr"]) / (a["dur"] * 1e-6) / 1e9 if a["dur"] else 0.0
    mean = lambda v: sum(v) / len(v) if v else float("nan")
    print("| %s | %d | %.1f | %.2f (%.2f + %.2f) | %.0f | %.3f | %.1f | %.1f | %.1f | %s | %s x %s | %.1f |"
          % (k, n, us, mb, a["rd"] / n / 1e6, a["wr"] / n / 1e6, gbs, gbs / peak, mean(a["occ"]), mean(a["dram_pct"]), mean(a["sm_pct"]),
             int(a["regs"] or 0), int(a["grid"] or 0), int(a["block"] or 0), mean(a["l1"])))
    out[k] = {"launches": n, "us_per_launch": us, "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "gbs": gbs}
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], "w"), indent=1)
