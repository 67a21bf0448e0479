#!/usr/bin/env python
"""One or two steps of a configuration between cudaProfilerStart/Stop, for ncu (DRAM evidence
for the streaming stages, not only the Gauss-Seidel kernels).

  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
      -o gpurun_out/r02_stages_batch python tools/profile_stages.py --scene batch
  ... --scene pile --n 100000 --at 14      (large-world mode 1, falling phase: LBVH rebuild + queries + add_pair)
  ... --scene mixed --n 10000 --at 120     (polygon narrowphase at config size)
  ... --scene add_pair --n 20000 --at 40

Nothing here is a benchmark: a run under ncu is never a bench value."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="batch", choices=["batch", "pile", "mixed", "add_pair"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--worlds", type=int, default=4096)
ap.add_argument("--at", type=int, default=400, help="untimed steps before the profiled ones")
ap.add_argument("--steps", type=int, default=1, help="profiled steps")
ap.add_argument("--mode", type=int, default=1, help="large-world mode for single-world scenes")
a = ap.parse_args()

import torch
from box2d_rs_b200 import scenes, sharding, world
from box2d_rs_b200.batch import Context

stream = torch.cuda.Stream()
ctx = Context(0, stream=stream.cuda_stream)
if a.scene == "batch":
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.pyramid(wg)
    wg.set_allow_sleeping(False)
    b = wg.batch(a.worlds, max_contacts=1024, solver="one_stream")  # full-size launches, no stream groups
    b.set_linear_velocity(211, sharding.perturbation(0, a.worlds, 0xB2D + 3))
    for _ in range(a.at // 100):
        b.step(scenes.DT, 8, 3, 100)
    ctx.sync()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        b.step(scenes.DT, 8, 3, 1)
    ctx.sync()
    torch.cuda.profiler.stop()
    st = b.stats()
    print("profiled %d step(s): %d worlds, %.1f contacts / %.1f touching per world, status %s"
          % (a.steps, a.worlds, st["contacts"].mean(), st["touching"].mean(), set(st["status"].tolist())))
else:
    gravity = (0.0, 0.0) if a.scene == "add_pair" else (0.0, -10.0)
    wg = world.B2world(gravity, ctx=ctx)
    getattr(scenes, a.scene)(wg, n=a.n or {"pile": 100000, "mixed": 10000, "add_pair": 20000}[a.scene])
    wg.set_large_mode(a.mode)
    for _ in range(a.at):
        wg.step(scenes.DT, 8, 3)
    ctx.sync()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        wg.step(scenes.DT, 8, 3)
    ctx.sync()
    torch.cuda.profiler.stop()
    st = wg.get_stats()
    print("profiled %d step(s) of %s: contacts %d touching %d islands %d moved %d created %d status %d"
          % (a.steps, a.scene, st["contacts"], st["touching"], st["islands"], st["moved"], st["created"], st["status"]))
