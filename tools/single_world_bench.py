#!/usr/bin/env python
"""ms/step of the single-large-world configurations of BASELINE.json (configs[1], [3], [4]) on one GPU, next
to the CPU oracle on the same scene (one host thread: the reference is single-threaded by construction).

  python tools/single_world_bench.py --scene pile --n 100000 --steps 30 [--check]

Honest caveat printed with the numbers: a single world has no batch parallelism; the ordered stages
(island DFS, Gauss-Seidel sweeps, tree re-insertion) run as one thread per world in exact reference order.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(kind, w, n):
    from box2d_rs_b200 import scenes
    if kind == "pile":
        scenes.pile(w, n=n)
    elif kind == "mixed":
        scenes.mixed(w, n=n)
    elif kind == "addpair":
        scenes.add_pair(w, n=n)
    else:
        raise SystemExit("unknown scene")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="pile", choices=["pile", "mixed", "addpair"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--skip", type=int, default=5, help="untimed leading steps")
    ap.add_argument("--check", action="store_true", help="compare the final state with the oracle bit for bit")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle (profiling runs)")
    ap.add_argument("--large", action="store_true",
                    help="large-world mode (b2gpu_world_set_large_mode): data-parallel broadphase / islands; --check then "
                         "verifies one teacher-forced step from the oracle's final state (contacts created in it as a set)")
    ap.add_argument("--large-exact", action="store_true",
                    help="large-world mode that keeps the replica tree (flag 2): reference contact order; --check compares the "
                         "free-running final state with the oracle bit for bit")
    args = ap.parse_args()
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    gravity = (0.0, 0.0) if args.scene == "addpair" else (0.0, -10.0)
    out = {"scene": args.scene, "bodies": args.n, "steps": args.steps, "dt": "1/60", "iters": "8/3", "allow_sleep": True}
    wo = b2o.B2world(gravity)
    t0 = time.time()
    build(args.scene, wo, args.n)
    out["build_s_oracle"] = time.time() - t0
    prof = {}
    t_cpu = 0.0
    for i in range(0 if args.no_cpu else args.steps):
        t0 = time.perf_counter()
        wo.step(scenes.DT, 8, 3)
        dt = time.perf_counter() - t0
        if i >= args.skip:
            t_cpu += dt
            for k, v in wo.get_profile().items():
                prof[k] = prof.get(k, 0.0) + v
    n_t = max(args.steps - args.skip, 1)
    out["cpu_ms_per_step"] = 1e3 * t_cpu / n_t
    out["cpu_profile_ms"] = {k: v / n_t for k, v in prof.items()}
    st = wo.get_stats()
    out["contacts"] = int(st["contacts"])
    out["touching"] = int(st["touching"])
    out["islands"] = int(st["islands"])
    if not args.no_gpu:
        wg = world.B2world(gravity)
        t0 = time.time()
        build(args.scene, wg, args.n)
        out["build_s_gpu_host_mirror"] = time.time() - t0
        out["mode"] = "large" if args.large else "large_exact" if args.large_exact else "exact"
        if args.large:
            wg.set_large_mode(1)
        elif args.large_exact:
            wg.set_large_mode(2)
        wg.ctx.set_profiling(True)
        t_gpu = 0.0
        for i in range(args.steps):
            if i == args.skip:
                wg.ctx.sync()
                wg.ctx.set_profiling(True)
            t0 = time.perf_counter()
            wg.step(scenes.DT, 8, 3)
            wg.ctx.sync()
            if i >= args.skip:
                t_gpu += time.perf_counter() - t0
        stages = wg.ctx.stage_times()
        out["gpu_ms_per_step"] = 1e3 * t_gpu / n_t
        out["gpu_stage_ms"] = {k: v[0] / n_t for k, v in stages.items() if v[1] > 0}
        out["gpu_over_cpu"] = out["cpu_ms_per_step"] / out["gpu_ms_per_step"] if not args.no_cpu else None
        gs = wg.get_stats()
        out["gpu_status"] = int(gs["status"])
        if args.check and args.large:
            import parity
            wg.upload(wo.snapshot())
            wg.set_large_mode(True)
            wo.step(scenes.DT, 8, 3)
            wg.step(scenes.DT, 8, 3)
            bad = parity.compare_large_step(wo.snapshot(), wg.snapshot(), wo.get_stats(), wg.get_stats())
            out["teacher_forced_step_identical_to_oracle"] = not bad
            if bad:
                out["mismatch"] = bad[:4]
        elif args.check:
            import parity
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot())
            bad += [b for b in parity.compare_stats(wo.get_stats(), wg.get_stats()) if "island_bodies" not in b or not args.large_exact]
            out["bit_identical_to_oracle"] = not bad
            if bad:
                out["mismatch"] = bad[:4]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
