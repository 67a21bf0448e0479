#!/usr/bin/env python
"""Jointed agents at batch scale (not the headline: bench.py measures BASELINE's Pyramid batch).  N worlds of the testbed car
(`scenes.car`: chassis on two wheel joints with springs, limits and a motor) or of another jointed scene, every world with its
own motor speed per step — the action of an RL policy — through b2gpu_batch_set_joint_control:

  * `device_ms_per_step`: the step alone, CUDA events on the launching stream, state resident;
  * `loop_ms_per_step`: action in (H2D + scatter), one step, body state out (gather + D2H), per step, wall clock between syncs;
  * `cpu_thread_ms_per_world_step`: the C++ oracle on one host thread;
  * worlds 0 and N-1 are compared with oracle worlds driven by the same actions, bit for bit, at the end.

    python tools/joint_batch_bench.py --worlds 4096 --steps 200 > profiles/r02_joint_batch.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="car", choices=["car", "joints_mix", "tumbler"])
ap.add_argument("--worlds", type=int, default=4096)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warmup", type=int, default=60)
a = ap.parse_args()

import numpy as np
import torch
import parity
from box2d_rs_b200 import scenes, world
from box2d_rs_b200.batch import Context
from oracle import b2o

stream = torch.cuda.Stream()
ctx = Context(0, stream=stream.cuda_stream)
recipe = {"car": lambda w: scenes.car(w)[0], "joints_mix": scenes.joints_mix, "tumbler": lambda w: scenes.tumbler(w, n=120)}[a.scene]
wg = world.B2world((0.0, -10.0), ctx=ctx)
joint = recipe(wg).index
wo = b2o.B2world((0.0, -10.0))
recipe(wo)
n = a.worlds
bt = wg.batch(n)
rng = np.random.default_rng(11)
total = a.warmup + 2 * a.steps
actions = rng.uniform(-25.0, 5.0, (total // 10 + 1, n)).astype(np.float32)  # a new motor speed per world every 10 steps


def act(i):
    if i % 10 == 0:
        bt.set_motor_speeds(joint, actions[i // 10])


i = 0
for _ in range(a.warmup):
    act(i); bt.step(scenes.DT, 8, 3); i += 1
ctx.sync()
# (a) the step alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
dev_ms = 0.0
for _ in range(a.steps):
    act(i)
    with torch.cuda.stream(stream):
        e0.record(stream)
        bt.step(scenes.DT, 8, 3)
        e1.record(stream)
    torch.cuda.synchronize()
    dev_ms += e0.elapsed_time(e1)
    i += 1
# (b) the loop: action in, step, observation out
t0 = time.perf_counter()
for _ in range(a.steps):
    act(i)
    bt.step(scenes.DT, 8, 3)
    state = bt.body_state()
    i += 1
ctx.sync()
loop_s = time.perf_counter() - t0
bad = []
st = bt.stats()
# the oracle through the same actions for two of the worlds, timing one host thread
orc = {w: wo.clone() for w in (0, n - 1)}
t0 = time.perf_counter()
for k in range(total):
    for w, o in orc.items():
        if k % 10 == 0:
            o.joint(joint).set_motor_speed(float(actions[k // 10][w]))
        o.step(scenes.DT, 8, 3)
cpu_s = time.perf_counter() - t0
for w, o in orc.items():
    bad += parity.compare_snapshots(o.snapshot(), bt.download_world(w))
nb = wo.get_body_count()
out = {
    "what": "jointed agents at batch scale (tools/joint_batch_bench.py), one B200; not the headline metric",
    "scene": a.scene, "worlds": n, "bodies_per_world": nb, "joints_per_world": wo.get_joint_count(),
    "action": "a new motor speed per world every 10 steps through b2gpu_batch_set_joint_control",
    "steps": a.steps, "warmup": a.warmup,
    "device_ms_per_step": dev_ms / a.steps,
    "device_world_steps_per_s": n * a.steps / (dev_ms * 1e-3),
    "loop_ms_per_step": 1e3 * loop_s / a.steps,
    "loop_world_steps_per_s": n * a.steps / loop_s,
    "loop_io": "per step: get_body_state D2H of %d bytes; every 10th step an action H2D of %d bytes" % (state.nbytes, 4 * n),
    "cpu_thread_ms_per_world_step": 1e3 * cpu_s / (total * len(orc)),
    "speedup_vs_one_cpu_thread": (n * a.steps / (dev_ms * 1e-3)) / (total * len(orc) / cpu_s),
    "contacts_per_world": float(st["contacts"].mean()), "status": sorted(set(int(x) for x in st["status"])),
    "oracle_checked_worlds": sorted(orc), "bit_identical_to_oracle": bad == [], "mismatches": bad[:4],
}
print(json.dumps(out))
