#!/usr/bin/env python
"""Throughput of the batched closest-hit ray cast (b2gpu_batch_ray_cast_closest / b2gpu_world_ray_cast_closest)
through HOST buffers (H2D of the rays and D2H of the hits inside the timed region), next to the oracle's
restatement of B2world::ray_cast on one host thread.

  python tools/ray_bench.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    out = {}
    rng = np.random.default_rng(1)
    # (a) RL-style range sensor: 4096 settled Pyramid worlds x 64 rays fanned from a point above the stack
    wg = world.B2world((0.0, -10.0))
    scenes.pyramid(wg)
    wo = b2o.B2world((0.0, -10.0))
    scenes.pyramid(wo)
    n_worlds, n_rays = 4096, 64
    bt = wg.batch(n_worlds, max_contacts=1024)
    bt.step(scenes.DT, 8, 3, 120)
    for _ in range(120):
        wo.step(scenes.DT, 8, 3)
    ang = np.linspace(-np.pi, 0.0, n_rays, dtype=np.float32)
    rays1 = np.stack([np.zeros(n_rays), np.full(n_rays, 30.0), 40.0 * np.cos(ang), 30.0 + 40.0 * np.sin(ang)], 1).astype(np.float32)
    rays = np.ascontiguousarray(np.broadcast_to(rays1, (n_worlds, n_rays, 4)))
    bt.ray_cast_closest(rays)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        got = bt.ray_cast_closest(rays)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    ref = wo.ray_cast_closest(rays1)
    dt_cpu = time.perf_counter() - t0
    same = bool(np.array_equal(ref["fraction"].view(np.uint32), got[7]["fraction"].view(np.uint32)) and
                np.array_equal(ref["fixture"], got[7]["fixture"]))
    out["pyramid_batch"] = {"worlds": n_worlds, "rays_per_world": n_rays, "ms_per_call": 1e3 * dt, "rays_per_s": n_worlds * n_rays / dt,
                            "cpu_rays_per_s_1_thread": n_rays / dt_cpu, "hits": int((got["fixture"] >= 0).sum()),
                            "world_7_equals_oracle": same}
    bt.close()
    wg.close()
    # (b) one world: the terrain scene, 200k random rays
    wg = world.B2world((0.0, -10.0))
    scenes.terrain(wg)
    wo = b2o.B2world((0.0, -10.0))
    scenes.terrain(wo)
    for _ in range(200):
        wg.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
    n = 200000
    p1 = rng.uniform((-120, -6), (120, 20), (n, 2))
    a = rng.uniform(0, 2 * np.pi, n)
    ln = rng.uniform(1.0, 60.0, n)
    rays = np.concatenate([p1, p1 + np.stack([np.cos(a), np.sin(a)], 1) * ln[:, None]], 1).astype(np.float32)
    wg.ray_cast_closest(rays[:1000])
    t0 = time.perf_counter()
    got = wg.ray_cast_closest(rays)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = wo.ray_cast_closest(rays[:20000])
    dt_cpu = time.perf_counter() - t0
    same = bool(np.array_equal(ref["fraction"].view(np.uint32), got[:20000]["fraction"].view(np.uint32)) and
                np.array_equal(ref["fixture"], got[:20000]["fixture"]))
    out["terrain_world"] = {"rays": n, "ms_per_call": 1e3 * dt, "rays_per_s": n / dt, "cpu_rays_per_s_1_thread": 20000 / dt_cpu,
                            "hits": int((got["fixture"] >= 0).sum()), "first_20000_equal_oracle": same}
    wg.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
