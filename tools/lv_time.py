#!/usr/bin/env python
"""Quick timing of the large-world solver stages of one scene (development aid for b2g_levels.h): runs the scene to --at,
then times --steps steps with the per-stage events on.  Not a benchmark."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="mixed", choices=["pile", "mixed", "add_pair"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--at", type=int, default=220)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--threshold", type=int, default=0)
ap.add_argument("--mode", type=int, default=1, help="large-world mode: 1 = LBVH order, 2 = replica tree kept (reference contact order)")
a = ap.parse_args()
from box2d_rs_b200 import scenes, world
from box2d_rs_b200.batch import Context
ctx = Context(0, lib_path=os.environ.get("B2GPU_LIB"))
gravity = (0.0, 0.0) if a.scene == "add_pair" else (0.0, -10.0)
wg = world.B2world(gravity, ctx=ctx)
getattr(scenes, a.scene)(wg, n=a.n or {"pile": 100000, "mixed": 10000, "add_pair": 20000}[a.scene])
wg.set_large_mode(a.mode)
if a.threshold:
    wg.set_level_threshold(a.threshold)
for _ in range(a.at):
    wg.step(scenes.DT, 8, 3)
ctx.sync()
ctx.set_profiling(True)
t0 = time.perf_counter()
for _ in range(a.steps):
    wg.step(scenes.DT, 8, 3)
ctx.sync()
dt = (time.perf_counter() - t0) / a.steps * 1e3
ms = {k: v[0] for k, v in ctx.stage_times().items()}
st = wg.get_stats()
print("direct-path visits (debug builds) %s" % (st["reserved"],)); print("%s at %d: %.2f ms/step; levels %d islands %d island_contacts %d; %s" % (
    a.scene, a.at, dt, st["solver_levels"], st["islands"], st["island_contacts"],
    {k: round(v / a.steps, 3) for k, v in ms.items() if v / a.steps > 0.05}))
