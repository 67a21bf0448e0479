#!/usr/bin/env python
"""Summarise an `ncu --metrics ... --csv --log-file X.csv` capture (long format: one row per launch and metric) per
kernel: duration, DRAM bytes, achieved GB/s, occupancy, registers, coalescing.

  python tools/ncu_table.py gpurun_out/r02_stages_batch.csv [peak_gbs] [out.json]
"""
import csv
import json
import re
import sys

peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6551.7
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)


def to_base(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    exact = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    if u in exact:  # durations -> microseconds
        return v * exact[u]
    for k, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * f
    return v


launch = {}
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"], "block": r["Block Size"], "grid": r["Grid Size"]})
    try:
        d[r["Metric Name"]] = to_base(r["Metric Value"], r["Metric Unit"])
    except ValueError:
        pass
agg = {}
for d in launch.values():
    m = re.search(r"stage_kernel(?:_occ)?<(?:b2g::)?(\w+)>", d["name"])
    short = m.group(1) if m else re.sub(r"\(.*", "", d["name"]).replace("void ", "").split("::")[-1][:60]
    a = agg.setdefault(short, {"n": 0, "dur": 0.0, "rd": 0.0, "wr": 0.0, "inst": 0.0, "occ": [], "dram": [], "sm": [], "l1": [],
                               "regs": 0, "grid": d["grid"], "block": d["block"], "maxgrid": 0})
    a["n"] += 1
    a["dur"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    a["inst"] += d.get("smsp__inst_executed.sum", 0.0)
    a["regs"] = max(a["regs"], d.get("launch__registers_per_thread", 0))
    g = d.get("launch__grid_size", 0)
    if g >= a["maxgrid"]:
        a["maxgrid"], a["grid"] = g, d["grid"]
    for k, dst in (("sm__warps_active.avg.pct_of_peak_sustained_active", "occ"),
                   ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram"),
                   ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm"),
                   ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "l1")):
        if k in d:
            a[dst].append((d[k], d.get("gpu__time_duration.sum", 1.0)))
total = sum(a["dur"] for a in agg.values())
print("| kernel | launches | us total | share | us / launch | DRAM MB / launch (rd + wr) | GB/s | of %.0f | warps active %% | dram %% | sm %% | regs | largest grid x block | L1 sectors / ld request |" % peak)
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
out = {}


def wmean(v):
    w = sum(x[1] for x in v)
    return sum(x[0] * x[1] for x in v) / w if w else float("nan")


for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["dur"]):
    n = a["n"]
    gbs = (a["rd"] + a["wr"]) / (a["dur"] * 1e-6) / 1e9 if a["dur"] else 0.0
    print("| %s | %d | %.1f | %.1f %% | %.1f | %.3f (%.3f + %.3f) | %.0f | %.3f | %.1f | %.1f | %.1f | %d | %s x %s | %.1f |"
          % (k, n, a["dur"], 100.0 * a["dur"] / total, a["dur"] / n, (a["rd"] + a["wr"]) / n / 1e6, a["rd"] / n / 1e6, a["wr"] / n / 1e6,
             gbs, gbs / peak, wmean(a["occ"]), wmean(a["dram"]), wmean(a["sm"]), int(a["regs"]), a["grid"], a["block"], wmean(a["l1"])))
    out[k] = {"launches": n, "us_total": a["dur"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "gbs": gbs,
              "warp_instructions": a["inst"]}
print("\ntotal %.1f us over %d launches" % (total, sum(a["n"] for a in agg.values())))
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], "w"), indent=1)
