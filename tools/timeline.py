"""Reads the Gauss-Seidel CTA timeline written with B2GPU_TIMELINE=<file> (see b2g_solver_smem.cuh) and prints,
for the last launches, per kernel launch: CTAs, first start, last end, median CTA duration (microseconds)."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 3)
kind = (a[:, 0] >> np.uint64(32)).astype(int)
cta = (a[:, 0] & np.uint64(0xff)).astype(int)
wb_first = ((a[:, 0] & np.uint64(0xffffffff)) >> np.uint64(8)).astype(int)
t0 = a[:, 1].astype(np.int64)
t1 = a[:, 2].astype(np.int64)
last = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
sel = slice(max(0, len(a) - last), len(a))
kind, wb_first, t0, t1 = kind[sel], wb_first[sel], t0[sel], t1[sel]
base = t0.min()
# a launch = (kind, wb_first) cluster in time: split where the start jumps by more than 100 us within the same key
rows = []
for key in sorted(set(zip(kind.tolist(), wb_first.tolist()))):
    m = (kind == key[0]) & (wb_first == key[1])
    s0, s1 = t0[m], t1[m]
    order = np.argsort(s0)
    s0, s1 = s0[order], s1[order]
    cuts = np.where(np.diff(s0) > 100_000)[0] + 1
    for seg0, seg1 in zip(np.split(s0, cuts), np.split(s1, cuts)):
        rows.append((seg0.min() - base, key[0], key[1], len(seg0), seg1.max() - base, float(np.median(seg1 - seg0)), float((seg1 - seg0).max())))
rows.sort()
print("start_us kernel wb_first ctas end_us median_cta_us max_cta_us")
for r in rows:
    print("%9.1f %s %4d %4d %9.1f %8.1f %8.1f" % (r[0] / 1e3, "vel" if r[1] == 1 else "pos", r[2], r[3], r[4] / 1e3, r[5] / 1e3, r[6] / 1e3))
