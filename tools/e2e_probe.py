#!/usr/bin/env python
"""Where the end-to-end call (bench.py `e2e`: one step per call through host buffers) spends its time over the device-resident
step: the same 4096 settled Pyramid worlds, one step per call, with the forces copy and / or the state copy left out.
Development aid, not a benchmark (wall clock between synchronisations, 100 calls each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from box2d_rs_b200 import scenes, sharding, world
from box2d_rs_b200.batch import Context
from box2d_rs_b200.lib import check
stream = torch.cuda.Stream()
ctx = Context(0, stream=stream.cuda_stream)
wg = world.B2world((0.0, -10.0), ctx=ctx)
scenes.pyramid(wg)
wg.set_allow_sleeping(False)
n = 4096
bt = wg.batch(n, max_contacts=1024)
bt.set_linear_velocity(211, sharding.perturbation(0, n, 0xB2D + 3))
for _ in range(4):
    bt.step(scenes.DT, 8, 3, 100)
ctx.sync()
nd = len(bt.dynamic_bodies())
forces = torch.zeros((n, nd, 3), dtype=torch.float32).pin_memory()
state = torch.zeros((n, nd, 6), dtype=torch.float32).pin_memory()
fp, sp = forces.numpy().ctypes.data, state.numpy().ctypes.data
L = bt.L
def run(f, s, calls=100):
    for _ in range(5):
        check(L, L.b2gpu_batch_step_host_dynamic(bt.h, f, s, scenes.DT, 8, 3, 1))
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(calls):
        check(L, L.b2gpu_batch_step_host_dynamic(bt.h, f, s, scenes.DT, 8, 3, 1))
    ctx.sync()
    return 1e3 * (time.perf_counter() - t0) / calls
def run_plain(per, calls):
    bt.step(scenes.DT, 8, 3, per); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(calls):
        bt.step(scenes.DT, 8, 3, per)
        if per == 1: ctx.sync()
    ctx.sync()
    return 1e3 * (time.perf_counter() - t0) / (calls * per)
print("ms per step: 20 steps per call %.3f | 1 step per call + sync %.3f | + forces in %.3f | + state out %.3f | both (= e2e) %.3f"
      % (run_plain(20, 5), run_plain(1, 100), run(fp, None), run(None, sp), run(fp, sp)))
