#!/usr/bin/env python
"""Dependency-DAG depth of the exact-order Gauss-Seidel sweeps of the LARGEST island, measured with the CPU oracle
("prove or break the sequential-by-definition claim").

  python tools/dag_depth.py --scene pile --n 100000 --steps 140 --every 10 --out profiles/r02_dag_pile100k.json

depth       = longest chain of visits that share a movable body, over warm start + velocity sweeps (exact order)
makespan_P  = simulated time (in visits) of the chunked dataflow schedule of the giant-island solver with P workers
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2d_rs_b200 import scenes
from oracle import b2o

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="pile")
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--steps", type=int, default=140)
ap.add_argument("--every", type=int, default=10)
ap.add_argument("--handover", type=float, default=2.0)
ap.add_argument("--out", default=None)
a = ap.parse_args()
gravity = (0.0, 0.0) if a.scene == "add_pair" else (0.0, -10.0)
w = b2o.B2world(gravity)
getattr(scenes, a.scene)(w, n=a.n)
rows = []
for i in range(a.steps):
    collect = (i % a.every == a.every - 1) or i == a.steps - 1
    w.set_collect_dag(collect, a.handover)
    t0 = time.perf_counter()
    w.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
    ms = 1e3 * (time.perf_counter() - t0)
    if collect:
        d = w.dag_stats()
        st = w.get_stats()
        visits = d["contacts"] * d["sweeps"]
        row = {"step": i, "cpu_ms": round(ms, 1), "islands": int(st["islands"]), "touching": int(st["touching"]),
               "largest_island_contacts": int(d["contacts"]), "largest_island_bodies": int(d["bodies"]), "visits": int(visits),
               "depth": int(d["depth"]), "depth_one_sweep": int(d["depth_one_sweep"]),
               "parallelism": round(visits / max(d["depth"], 1), 1),
               "makespan": {k[9:]: int(d[k]) for k in d if k.startswith("makespan_")}}
        rows.append(row)
        print(json.dumps(row), flush=True)
if a.out:
    json.dump({"scene": a.scene, "n": a.n, "handover_visits": a.handover, "rows": rows}, open(a.out, "w"), indent=1)
