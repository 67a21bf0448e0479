#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list of the large-world mode
(profiles/r01_large_launches.csv): microseconds and launches per step, over the last N steps of the capture
(a step starts at LwStatsResetK).  Times under ncu are cold-cache and serialised: read the SHARES.

  python tools/launch_summary.py profiles/r01_large_launches.csv [--steps 4]
"""
import argparse
import collections
import csv
import re


def short(name):
    m = re.search(r"stage_kernel(?:_occ)?<b2g::(\w+)>", name)
    if m:
        return m.group(1)
    m = re.search(r"(DeviceRadixSort\w+|DeviceScan\w+)", name)
    return m.group(1) if m else name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    hdr, data = None, []
    for r in csv.reader(open(args.csv)):
        if r and r[0] == "ID":
            hdr = r
        elif hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    names = [short(d["Kernel Name"]) for d in data]
    starts = [i for i, n in enumerate(names) if n == "LwStatsResetK"]
    if len(starts) <= args.steps:
        raise SystemExit("capture holds %d steps only" % len(starts))
    lo, hi = starts[-args.steps - 1], starts[-1]
    tot, cnt = collections.Counter(), collections.Counter()
    for d, n in zip(data[lo:hi], names[lo:hi]):
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else v * 1e3 if unit.startswith("m") else v
        tot[n] += v
        cnt[n] += 1
    print("%d launches captured, %d steps; last %d steps: %.1f us and %.1f launches per step"
          % (len(data), len(starts), args.steps, sum(tot.values()) / args.steps, sum(cnt.values()) / args.steps))
    for n, v in tot.most_common():
        print("%-36s %8.1f us  x%.1f" % (n, v / args.steps, cnt[n] / args.steps))


if __name__ == "__main__":
    main()
