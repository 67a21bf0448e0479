#!/usr/bin/env python
"""bench.py — body-steps/s of the B2world::step hot path on batched Pyramid worlds (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--worlds 4096]

A "step" is one B2world::step (dt = 1/60, 8 velocity / 3 position iterations, continuous off) of every
world of the batch.  Workload = BASELINE.json configs[2]: 4096 independent testbed-Pyramid worlds
(212 bodies, 210 dynamic boxes each) per GPU, each perturbed by a seeded random initial velocity of
its top box, settled by an untimed pre-roll, sleeping disabled in both engines so every timed step
does the same work.  body-steps = 210 dynamic bodies x worlds x steps.

ours:       value  = device-resident throughput (CUDA events, max over ranks)
            e2e    = the same steps through b2gpu_batch_step_host with pinned HOST buffers: per step
                     H2D of per-body forces and D2H of the body state are inside the timed region
reference:  the reference's CPU path (the C++ oracle restating box2d-rs; the Rust crate cannot be
            built here) with one world per host thread on all host cores, on a bounded sample.
Multi-GPU (torchrun): worlds are sharded by world, no traffic during the step ("weak": 4096 worlds per
GPU); one NCCL all_gather of per-rank state digests after the timed region validates the shards.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DYNAMIC_BODIES = 210
PREROLL = 400
SEED = 0xB2D + 3


def perturbation(n_worlds, rank):
    """Initial velocity of each world's top box: a function of the global world index (sharding.py)."""
    from box2d_rs_b200 import sharding
    return sharding.perturbation(rank * n_worlds, n_worlds, SEED)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.interval = index, [], False, 0.2

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


def cpu_arm(n_sample_worlds, inner_steps, threads, repeats, warmup):
    """Reference CPU path on a bounded sample: returns (body-steps/s, seconds per repeat list)."""
    from box2d_rs_b200 import scenes
    from oracle import b2o  # the CPU restatement: allowed here as the measured reference arm only
    proto = b2o.B2world((0.0, -10.0))
    scenes.pyramid(proto)
    proto.set_allow_sleeping(False)
    v = perturbation(n_sample_worlds, 0)
    worlds = []
    for i in range(n_sample_worlds):
        w = proto.clone()
        w.body(211).set_linear_velocity((float(v[i, 0]), float(v[i, 1])))
        worlds.append(w)
    b2o.run_worlds_mt(worlds, PREROLL, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)
    for _ in range(warmup):
        b2o.run_worlds_mt(worlds, inner_steps, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)
    secs = [b2o.run_worlds_mt(worlds, inner_steps, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)
            for _ in range(repeats)]
    total = DYNAMIC_BODIES * n_sample_worlds * inner_steps * repeats
    return total / sum(secs), secs


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import b2o
    threads = b2o.hardware_threads() or os.cpu_count() or 1
    n_sample = max(2 * threads, 8)
    inner = 25
    value, secs = cpu_arm(n_sample, inner, threads, args.steps, args.warmup)
    ms = 1e3 * sum(secs) / len(secs)
    line = {
        "impl": "reference", "metric": "body-steps/s, batched Pyramid worlds", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "4096 batched independent Pyramid worlds (BASELINE configs[2])",
                   "sample": "%d worlds x %d world-steps per bench step, one world per host thread" % (n_sample, inner),
                   "dt": "1/60", "velocity_iterations": 8, "position_iterations": 3, "allow_sleep": False,
                   "state": "settled (%d pre-roll steps), top box perturbed per world" % PREROLL},
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d worlds x %d steps x %d repeats (C++ oracle restating box2d-rs; the Rust crate "
                                   "cannot be built in this image)" % (n_sample, inner, args.steps)},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    OUT.emit(json.dumps(line))


def single_world_leg(ctx):
    """The second half of BASELINE.json's metric: ms/step of ONE large world (configs[3] pile-100k, configs[4]
    AddPair-20k) in the large-world mode (b2gpu_world_set_large_mode), next to the oracle on one host thread
    (the reference steps a world on one thread by construction).  Bounded: a few dozen steps from t = 0, both
    engines over the same steps.  Wall clock around step + sync: the mode is host-driven (scalar readbacks)."""
    from box2d_rs_b200 import scenes, world
    from oracle import b2o  # measured CPU arm
    out = {"mode": "large-world (data-parallel broadphase / destruction / islands; exact Gauss-Seidel order per island)",
           "cpu": "C++ oracle restating box2d-rs, 1 thread",
           "note": "free-running: contacts created in one update_pairs call are appended in LBVH order, so after many steps the "
                   "trajectory (and the contact counts) differ from the oracle's, as two valid Box2D runs do; every single step "
                   "is the oracle's step of the same state (tests: test_large_mode_teacher_forced); "
                   "b2gpu_world_set_large_mode(w, 2) keeps the reference order at the price of sequential tree updates"}
    cases = [("addpair20k", lambda w: scenes.add_pair(w, n=20000), (0.0, 0.0), 5, 40),
             ("pile100k", lambda w: scenes.pile(w, n=100000), (0.0, -10.0), 3, 12)]
    for name, recipe, gravity, skip, steps in cases:
        try:
            wo = b2o.B2world(gravity)
            recipe(wo)
            t_cpu = 0.0
            for i in range(steps):
                t0 = time.perf_counter()
                wo.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
                if i >= skip:
                    t_cpu += time.perf_counter() - t0
            so = wo.get_stats()
            del wo
            wg = world.B2world(gravity, ctx=ctx)
            recipe(wg)
            wg.set_large_mode(True)
            t_gpu = 0.0
            for i in range(steps):
                t0 = time.perf_counter()
                wg.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
                ctx.sync()
                if i >= skip:
                    t_gpu += time.perf_counter() - t0
            sg = wg.get_stats()
            n_prof = 4  # stage split from a few further steps with an event pair around every launch (not timed above)
            ctx.set_profiling(True)
            for _ in range(n_prof):
                wg.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
            stages = ctx.stage_times()
            ctx.set_profiling(False)
            n_t = steps - skip
            out[name] = {"ms_per_step": 1e3 * t_gpu / n_t, "cpu_ms_per_step": 1e3 * t_cpu / n_t,
                         "speedup_vs_cpu_thread": t_cpu / t_gpu, "steps": "%d-%d from t=0" % (skip, steps - 1),
                         "contacts": int(sg["contacts"]), "touching": int(sg["touching"]), "islands": int(sg["islands"]),
                         "status": int(sg["status"]),
                         "oracle_contacts": int(so["contacts"]), "oracle_touching": int(so["touching"]),
                         "stage_ms_next_%d_steps" % n_prof: {k: round(v[0] / n_prof, 4) for k, v in stages.items() if v[1] > 0}}
            wg.close()
        except Exception as e:  # never lose the main line over the secondary leg
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            try:
                ctx.set_profiling(False)
            except Exception:
                pass
    return out


def run_ours(args, rank, world_size, local_rank):
    import torch
    from box2d_rs_b200 import scenes, world
    from box2d_rs_b200.batch import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = Context(local_rank, stream=stream.cuda_stream)
    n_worlds = args.worlds
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.pyramid(wg)
    wg.set_allow_sleeping(False)
    batch = wg.batch(n_worlds, max_contacts=args.max_contacts, solver=args.solver)
    batch.set_linear_velocity(211, perturbation(n_worlds, rank))
    batch.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, PREROLL)
    ctx.sync()
    st = batch.stats()
    if (st["status"] != 0).any():
        raise SystemExit("bench.py: device status %s after pre-roll" % set(st["status"].tolist()))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)  # samples while the device is under this load (warm-up + timed region)
    sampler.start()
    batch.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, max(args.warmup, 300))  # ~0.5 s under load for the clock sampler
    # one untimed call with the timed call's signature: the library captures a CUDA graph per signature
    batch.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, args.steps)
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        batch.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, args.steps)
        e1.record(stream)
    barrier()
    clocks = sampler.summary()
    launches = ctx.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = DYNAMIC_BODIES * n_worlds * world_size * args.steps / (ms_total * 1e-3)

    # ---- per-stage device times (CUDA events around every launch on the launching stream)
    ctx.set_profiling(True)
    batch.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, args.steps)
    stages = ctx.stage_times()
    ctx.set_profiling(False)
    alg_bytes_step = batch.algorithmic_bytes()
    st = batch.stats()
    stage_ms = {k: v[0] / args.steps for k, v in stages.items() if v[1] > 0}
    top = max(stage_ms, key=stage_ms.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # algorithmic bytes of the dominant stage (DESIGN.md "Stages"): the ordered velocity stage reads each
    # island contact's 160 B constraint record and writes its 16 B impulses once, and reads+writes 24 B of
    # velocity per island body; every Gauss-Seidel iteration beyond that is on-chip in the roofline model.
    isl_contacts = int(st["island_contacts"].sum())
    isl_bodies = int(st["island_bodies"].sum())
    stage_alg = {"velocity": 176 * isl_contacts + 48 * isl_bodies, "position": 144 * isl_contacts + 56 * isl_bodies}
    top_alg = stage_alg.get(top, alg_bytes_step)
    achieved = top_alg / (stage_ms[top] * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this configuration
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tr.get("worlds") == n_worlds:
            traffic = tr.get(top)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": top_alg,
                "kernel_ms_per_launch": stage_ms[top],
                "whole_step": {"algorithmic_bytes": alg_bytes_step, "achieved": alg_bytes_step / (ms_per_step * 1e-3) / 1e9,
                               "frac": alg_bytes_step / (ms_per_step * 1e-3) / 1e9 / peak},
                "stage_ms": stage_ms}

    # ---- end to end through host buffers (pinned): forces H2D + step + state D2H per step
    nb = batch.body_count
    forces = torch.zeros((n_worlds, nb, 3), dtype=torch.float32).pin_memory()
    state = torch.zeros((n_worlds, nb, 8), dtype=torch.float32).pin_memory()
    f_np, s_np = forces.numpy(), state.numpy()
    for _ in range(max(args.warmup, 3)):
        batch.step_host(f_np, s_np, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.step_host(f_np, s_np, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, 1)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": DYNAMIC_BODIES * n_worlds * world_size * args.steps / e2e_s, "unit": "body-steps/s",
           "h2d_bytes_per_step": int(forces.numel() * 4), "d2h_bytes_per_step": int(state.numel() * 4),
           "ms_per_step": 1e3 * e2e_s / args.steps}

    # ---- validation gather over NCCL (outside the timed regions): per-rank digest of the body state
    from box2d_rs_b200 import sharding
    digests = sharding.world_digests(batch.body_state())
    gathered = None
    if dist is not None:
        per_rank = sharding.allgather_digests(dist, digests, device="cuda")
        gathered = {"worlds": int(sum(len(d) for d in per_rank)), "finite": bool(all(np.isfinite(d).all() for d in per_rank)),
                    "per_rank_sum": [float(d.sum()) for d in per_rank]}

    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        from oracle import b2o
        threads = b2o.hardware_threads() or os.cpu_count() or 1
        n_sample = max(2 * threads, 8)
        inner = 400
        cv, secs = cpu_arm(n_sample, inner, threads, 3, 1)
        cpu_baseline = {"value": cv, "unit": "body-steps/s", "cores": threads, "kind": "port",
                        "sample": "%d Pyramid worlds x %d steps x 3 repeats, one world per host thread (C++ oracle "
                                  "restating box2d-rs; %.1f s of CPU work)" % (n_sample, inner, sum(secs) * threads)}
    single_world = None
    if rank == 0 and world_size == 1 and not args.no_single_world:
        batch.close()
        single_world = single_world_leg(ctx)
    if rank == 0:
        line = {
            "metric": "body-steps/s, batched Pyramid worlds", "value": value, "unit": "body-steps/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d batched independent Pyramid worlds per GPU (BASELINE configs[2])" % n_worlds,
                       "worlds_per_gpu": n_worlds, "bodies_per_world": 212, "dynamic_bodies_per_world": DYNAMIC_BODIES,
                       "dt": "1/60", "velocity_iterations": 8, "position_iterations": 3, "allow_sleep": False,
                       "continuous_physics": False,
                       "state": "settled (%d pre-roll steps), top box perturbed per world (seed %d)" % (PREROLL, SEED),
                       "l2": "inputs exceed L2: per-step state of the batch is %.0f MB" % (alg_bytes_step / 1e6),
                       "contacts_per_world": float(st["contacts"].mean()), "touching_per_world": float(st["touching"].mean()),
                       "parallelism": "worlds sharded by index, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "validation_allgather": gathered, "single_world": single_world,
        }
        OUT.emit(json.dumps(line))
    batch.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


class StdoutToStderr:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the JSON line is
    written to the real stdout through `emit`."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


OUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=4096, help="worlds per GPU")
    ap.add_argument("--max-contacts", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-single-world", action="store_true", help="skip the single-large-world leg (ms/step of one world)")
    ap.add_argument("--solver", default=None, help="diagnostic: lane | generic | levels (default: best measured)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__
    if rank == 0 or not os.path.exists(__graft_entry__.SO):
        try:
            __graft_entry__.build()
        except Exception as e:  # the prebuilt .so travels with the snapshot; rebuilding is best effort
            if not os.path.exists(__graft_entry__.SO):
                raise
            print("bench.py: build skipped (%s)" % e, file=sys.stderr)
    global OUT
    with StdoutToStderr() as OUT:
        if args.impl == "reference":
            run_reference(args, rank)
        else:
            run_ours(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
