#!/usr/bin/env python
"""bench.py — body-steps/s of the B2world::step hot path on batched Pyramid worlds (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--worlds 4096]

A "step" is one B2world::step (dt = 1/60, 8 velocity / 3 position iterations, continuous off) of every world
of the batch.  Workload = BASELINE.json configs[2]: 4096 independent testbed-Pyramid worlds IN TOTAL (212
bodies, 210 dynamic boxes each), sharded by contiguous world ranges over the N ranks (sharding.world_range:
SURVEY §8e, "strong" scaling: 4096 / N worlds per GPU).  Every world is perturbed by a seeded initial velocity
of its top box (a function of the global world index), sleeping is disabled in both engines so every timed step
does the same work, and the timed steps follow an untimed 400-step pre-roll (settled phase).
body-steps = 210 dynamic bodies x worlds x steps.

ours:       value        = device-resident throughput (CUDA events on the launching stream, max over ranks)
            e2e          = the same steps through b2gpu_batch_step_host_dynamic with pinned HOST buffers: per step the
                           H2D of per-body forces and the D2H of the body state are inside the timed region
            run_from_t0  = the reference's own case (configs[0]: 1000 steps from t = 0: falling, settling,
                           settled) for the same 4096 worlds, device-resident, CPU arm beside it
            weak_4096_per_gpu (N > 1 only) = round 1's workload: 4096 worlds on EVERY GPU
            single_world (N = 1 only) = ms/step of the single-large-world configs on SURVEY §8d's windows
reference:  the reference's CPU path (the C++ oracle restating box2d-rs; the Rust crate cannot be built here) with
            one world per host thread on all host cores (persistent pool), on a bounded sample.
Multi-GPU (torchrun): no traffic during the step; after the timed region one NCCL all_gather of per-world state
digests, and every rank checks sampled worlds of its shard against the oracle bit for bit.
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
PREROLL_CHUNK = 100
SEED = 0xB2D + 3
CPU_INNER = 400  # world-steps per world inside one timed CPU call


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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_worlds(first_world, count):
    """`count` perturbed Pyramid worlds of the oracle, global world indices first_world.."""
    from box2d_rs_b200 import scenes, sharding
    from oracle import b2o  # the CPU restatement: allowed here as the measured reference arm / checker only
    proto = b2o.B2world((0.0, -10.0))
    scenes.pyramid(proto)
    proto.set_allow_sleeping(False)
    v = sharding.perturbation(first_world, count, SEED)
    worlds = []
    for i in range(count):
        w = proto.clone()
        w.body(211).set_linear_velocity((float(v[i, 0]), float(v[i, 1])))
        worlds.append(w)
    return worlds


def cpu_arm(threads, repeats, warmup, inner=CPU_INNER, from_t0_steps=0):
    """Reference CPU path on a bounded sample (2 worlds per host thread, one world per thread at a time).
    Returns {value, secs, n_sample, inner, from_t0}: body-steps/s of the settled phase, and (optionally) of a
    whole run of `from_t0_steps` steps from t = 0."""
    from box2d_rs_b200 import scenes
    from oracle import b2o
    n_sample = max(2 * threads, 8)
    worlds = cpu_worlds(0, n_sample)
    out = {"n_sample": n_sample, "inner": inner, "from_t0": None}
    b2o.run_worlds_mt(cpu_worlds(0, threads), 2, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)  # starts the pool
    if from_t0_steps:
        fresh = cpu_worlds(0, n_sample)
        s = b2o.run_worlds_mt(fresh, from_t0_steps, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)
        out["from_t0"] = {"steps": from_t0_steps, "value": DYNAMIC_BODIES * n_sample * from_t0_steps / s, "seconds": s}
        del fresh
    b2o.run_worlds_mt(worlds, PREROLL, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)
    for _ in range(warmup):
        b2o.run_worlds_mt(worlds, inner, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads)
    secs = [b2o.run_worlds_mt(worlds, inner, scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS, threads) for _ in range(repeats)]
    out["secs"] = secs
    out["value"] = DYNAMIC_BODIES * n_sample * inner * repeats / sum(secs)
    return out


def workload_config(total_worlds, world_size, extra=None):
    cfg = {"workload": "%d batched independent Pyramid worlds in total (BASELINE configs[2]), sharded by world over %d GPU(s)"
                       % (total_worlds, world_size),
           "total_worlds": total_worlds, "bodies_per_world": 212, "dynamic_bodies_per_world": DYNAMIC_BODIES,
           "dt": "1/60", "velocity_iterations": 8, "position_iterations": 3, "allow_sleep": False, "continuous_physics": False,
           "state": "settled (%d pre-roll steps), top box perturbed per world (seed %d)" % (PREROLL, SEED)}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args, rank, world_size):
    if rank != 0:
        return
    from oracle import b2o
    threads = b2o.hardware_threads() or os.cpu_count() or 1
    r = cpu_arm(threads, max(args.steps, 1), max(args.warmup, 0))
    value, secs = r["value"], r["secs"]
    ms = 1e3 * sum(secs) / len(secs)
    sample = ("%d worlds x %d world-steps per bench step, one world per host thread, persistent pool of %d threads "
              "(C++ oracle restating box2d-rs; the Rust crate cannot be built in this image)" % (r["n_sample"], r["inner"], threads))
    line = {
        "impl": "reference", "metric": "body-steps/s, batched Pyramid worlds", "value": value, "unit": "body-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.worlds, world_size, {"sample": sample}),
        "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    OUT.emit(json.dumps(line))


# ------------------------------------------------------------------------------------------ single large world
SINGLE_WORLD_CASES = [
    # name, scene, n, gravity, first timed step, steps timed here, SURVEY 8d window, modes
    ("pile100k", "pile", 100000, (0.0, -10.0), 100, 8, "100-500 (settled pile)", (1,)),
    ("addpair20k", "add_pair", 20000, (0.0, 0.0), 20, 40, "0-600 (the plough crosses the cloud in steps 24-60)", (1, 2)),
    ("mixed10k", "mixed", 10000, (0.0, -10.0), 200, 40, "0-1000", (1, 2)),
]


def single_world_leg(ctx, full=False):
    """The second half of BASELINE.json's metric: ms/step of ONE large world in the large-world modes
    (b2gpu_world_set_large_mode 1 = LBVH contact-creation order, 2 = reference order via the replica tree), next to the
    oracle on one host thread (the reference steps a world on one thread by construction).  Bounded: a slice at the
    start of each SURVEY §8d window (tools/single_world_bench.py runs whole windows; results under profiles/).  Both
    engines run free from t = 0 and are timed over the same step numbers; wall clock around step + sync."""
    from box2d_rs_b200 import scenes, world
    from oracle import b2o  # measured CPU arm
    out = {"cpu": "C++ oracle restating box2d-rs, 1 thread",
           "modes": {"1": "large-world mode, contacts of one update_pairs call appended in LBVH order (every step is the "
                          "oracle's step of the same state; free-running trajectories are two valid Box2D runs)",
                     "2": "large-world mode keeping the reference's contact-creation order (sequential replica-tree "
                          "re-insertion); free-running state bit-identical to the oracle"},
           "limit": "inside one island the exact Gauss-Seidel order is a dependency DAG 3-9x wider than a chain "
                    "(profiles/r02_dag_depth.md); islands of >= 1024 contacts are swept level by level by one CTA "
                    "(b2g_levels.h, profiles/r02_levels.md: a level costs ~1.7 visit latencies), every other island by one thread"}
    for name, scene, n, gravity, first, count, window, modes in SINGLE_WORLD_CASES:
        if full:
            count = {"pile100k": 100, "addpair20k": 580, "mixed10k": 800}[name]
        try:
            recipe = getattr(scenes, scene)
            wo = b2o.B2world(gravity)
            recipe(wo, n=n)
            t_cpu = 0.0
            for i in range(first + count):
                t0 = time.perf_counter()
                wo.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
                if i >= first:
                    t_cpu += time.perf_counter() - t0
            so = wo.get_stats()
            del wo
            case = {"survey_window": window, "timed_steps": "%d-%d from t=0" % (first, first + count - 1),
                    "cpu_ms_per_step": 1e3 * t_cpu / count, "oracle_contacts": int(so["contacts"]),
                    "oracle_touching": int(so["touching"]), "oracle_islands": int(so["islands"])}
            for mode in modes:
                wg = world.B2world(gravity, ctx=ctx)
                recipe(wg, n=n)
                wg.set_large_mode(mode)
                t_gpu = 0.0
                for i in range(first + count):
                    t0 = time.perf_counter()
                    wg.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
                    ctx.sync()
                    if i >= first:
                        t_gpu += time.perf_counter() - t0
                sg = wg.get_stats()
                n_prof = 2  # stage split from two further steps with an event pair around every launch (not timed above)
                ctx.set_profiling(True)
                for _ in range(n_prof):
                    wg.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
                stages = ctx.stage_times()
                ctx.set_profiling(False)
                case["mode%d" % mode] = {
                    "ms_per_step": 1e3 * t_gpu / count, "speedup_vs_cpu_thread": t_cpu / t_gpu,
                    "contacts": int(sg["contacts"]), "touching": int(sg["touching"]), "islands": int(sg["islands"]),
                    "status": int(sg["status"]),
                    "stage_ms_next_%d_steps" % n_prof: {k: round(v[0] / n_prof, 4) for k, v in stages.items() if v[1] > 0}}
                wg.close()
            out[name] = case
        except Exception as e:  # never lose the main line over the secondary leg
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            try:
                ctx.set_profiling(False)
            except Exception:
                pass
    return out


# ------------------------------------------------------------------------------------------ our arm
class Arm:
    """One batch of perturbed Pyramid worlds on this rank's GPU plus the timing helpers."""

    def __init__(self, args, ctx, stream, dist, first_world, n_worlds):
        from box2d_rs_b200 import scenes, sharding, world
        self.args, self.ctx, self.stream, self.dist = args, ctx, stream, dist
        self.first_world, self.n_worlds = first_world, n_worlds
        self.scenes = scenes
        self.wg = world.B2world((0.0, -10.0), ctx=ctx)
        scenes.pyramid(self.wg)
        self.wg.set_allow_sleeping(False)
        self.proto = self.wg.snapshot()
        self.batch = self.wg.batch(n_worlds, max_contacts=args.max_contacts, solver=args.solver)
        self.perturb = sharding.perturbation(first_world, n_worlds, SEED)
        self.batch.set_linear_velocity(211, self.perturb)
        self.steps_done = 0

    def step(self, n):
        s = self.scenes
        self.batch.step(s.DT, s.VEL_ITERS, s.POS_ITERS, n)
        self.steps_done += n

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        if self.dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(self, calls, steps_per_call):
        """CUDA events on the launching stream around `calls` x step(steps_per_call); ms, max over ranks."""
        import torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(calls):
                self.step(steps_per_call)
            e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def preroll(self):
        for _ in range(PREROLL // PREROLL_CHUNK):
            self.step(PREROLL_CHUNK)
        self.ctx.sync()
        self.batch.check_status()

    def e2e(self, steps, warmup):
        import torch
        s = self.scenes
        nd = len(self.batch.dynamic_bodies())  # compact I/O: the 210 dynamic boxes (5040 B of state per world)
        forces = torch.zeros((self.n_worlds, nd, 3), dtype=torch.float32).pin_memory()
        state = torch.zeros((self.n_worlds, nd, 6), dtype=torch.float32).pin_memory()
        f_np, s_np = forces.numpy(), state.numpy()
        for _ in range(max(warmup, 3)):
            self.batch.step_host_dynamic(f_np, s_np, s.DT, s.VEL_ITERS, s.POS_ITERS, 1)
            self.steps_done += 1
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.batch.step_host_dynamic(f_np, s_np, s.DT, s.VEL_ITERS, s.POS_ITERS, 1)
        self.barrier()
        sec = self.max_over_ranks(time.perf_counter() - t0)
        self.steps_done += steps
        return sec, int(forces.numel() * 4), int(state.numel() * 4)

    def validate(self, n_sample=2):
        """Sampled worlds of this rank's shard against the oracle, bit for bit: same perturbation (a function of the global
        world index), same number of steps.  Returns (global indices, all equal?)."""
        from box2d_rs_b200 import scenes
        picks = sorted(set([0, self.n_worlds - 1][:n_sample]))
        state = self.batch.body_state()
        ok = True
        for p in picks:
            (w,) = cpu_worlds(self.first_world + p, 1)
            for _ in range(self.steps_done):
                w.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
            ok = ok and np.array_equal(w.body_state().view(np.uint32), state[p].view(np.uint32))
        return [self.first_world + p for p in picks], bool(ok)

    def close(self):
        self.batch.close()
        self.wg.close()


def run_ours(args, rank, world_size, local_rank):
    import torch
    from box2d_rs_b200 import scenes, sharding
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
    total = args.worlds
    first, end = sharding.world_range(total, rank, world_size)
    arm = Arm(args, ctx, stream, dist, first, end - first)
    arm.preroll()

    # ---- device-resident throughput (settled phase)
    sampler = ClockSampler(local_rank)  # samples while the device is under this load (warm-up + timed region)
    sampler.start()
    for _ in range(3):
        arm.step(PREROLL_CHUNK)  # ~0.5 s under load for the clock sampler
    for _ in range(max(args.warmup, 3)):
        arm.step(1)
    arm.step(args.steps)  # one untimed call with the timed call's signature: the library captures a CUDA graph per signature
    launches0 = ctx.launch_count()
    ms_total = arm.timed_device(1, args.steps)
    launches = ctx.launch_count() - launches0
    clocks = sampler.summary()
    ms_per_step = ms_total / args.steps
    value = DYNAMIC_BODIES * total * args.steps / (ms_total * 1e-3)

    # ---- per-stage device times (CUDA events around every launch on the launching stream)
    ctx.set_profiling(True)
    arm.step(args.steps)
    stages = ctx.stage_times()
    ctx.set_profiling(False)
    batch = arm.batch
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
    # algorithmic bytes of the dominant stage (DESIGN.md "Stages"): the ordered velocity stage reads each island
    # contact's 160 B constraint record and writes its 16 B impulses once, and reads+writes 24 B of velocity per
    # island body; every Gauss-Seidel iteration beyond that is on-chip in the roofline model.
    isl_contacts = int(st["island_contacts"].sum())
    isl_bodies = int(st["island_bodies"].sum())
    stage_alg = {"velocity": 176 * isl_contacts + 48 * isl_bodies, "position": 144 * isl_contacts + 56 * isl_bodies,
                 "solve": 320 * isl_contacts + 104 * isl_bodies}
    top_alg = stage_alg.get(top, alg_bytes_step)
    achieved = top_alg / (stage_ms[top] * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture of this configuration
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tr.get("worlds_per_gpu") == arm.n_worlds:
            traffic = tr.get(top)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": top_alg,
                "kernel_ms_per_launch": stage_ms[top],
                "note": "rank 0's shard; per-stage times are CUDA events around every launch with the stream groups serialised",
                "whole_step": {"algorithmic_bytes": alg_bytes_step, "achieved": alg_bytes_step / (ms_per_step * 1e-3) / 1e9,
                               "frac": alg_bytes_step / (ms_per_step * 1e-3) / 1e9 / peak},
                "stage_ms": stage_ms}

    # ---- end to end through host buffers (pinned): forces H2D + step + state D2H per step
    e2e_s, h2d, d2h = arm.e2e(args.steps, args.warmup)
    e2e = {"value": DYNAMIC_BODIES * total * args.steps / e2e_s, "unit": "body-steps/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
           "bytes_are": "per rank (this rank's shard of %d worlds)" % arm.n_worlds,
           "layout": "b2gpu_batch_step_host_dynamic: per dynamic body 3 floats of force in, (c, a, v, w) = 6 floats out"}

    # ---- validation: sampled worlds against the oracle (every rank), digests of all worlds gathered over NCCL
    picks, ok = arm.validate()
    digests = sharding.world_digests(batch.body_state())
    validation = {"oracle_checked_worlds": picks, "bit_identical_to_oracle": ok, "steps_compared": arm.steps_done}
    if dist is not None:
        per_rank = sharding.allgather_digests(dist, digests, device="cuda")
        flags = [None] * world_size
        dist.all_gather_object(flags, {"rank": rank, "worlds": picks, "ok": ok})
        validation = {"allgather_worlds": int(sum(len(d) for d in per_rank)),
                      "finite": bool(all(np.isfinite(d).all() for d in per_rank)),
                      "distinct_digests": int(len(set(np.concatenate(per_rank).tolist()))),
                      "oracle_check_per_rank": flags, "bit_identical_to_oracle": bool(all(f["ok"] for f in flags)),
                      "steps_compared": arm.steps_done}
    else:
        validation["distinct_digests"] = int(len(set(digests.tolist())))

    # ---- configs[0]-style whole run: 1000 steps from t = 0 (falling, settling, settled), device-resident
    run_t0 = None
    if not args.no_t0:
        batch.reset(arm.proto)
        batch.set_linear_velocity(211, arm.perturb)
        t0_steps = 1000
        ms_t0 = arm.timed_device(t0_steps // PREROLL_CHUNK, PREROLL_CHUNK)
        run_t0 = {"steps": t0_steps, "ms_total": ms_t0, "ms_per_step": ms_t0 / t0_steps,
                  "value": DYNAMIC_BODIES * total * t0_steps / (ms_t0 * 1e-3), "unit": "body-steps/s",
                  "what": "BASELINE configs[0] (Pyramid, 1000 steps from t = 0) for all %d worlds, sleeping off" % total}

    # ---- round 1's workload for comparison (N > 1): 4096 worlds on EVERY GPU
    weak = None
    if world_size > 1 and not args.no_weak:
        arm.close()
        per_gpu = args.worlds
        arm = Arm(args, ctx, stream, dist, rank * per_gpu, per_gpu)
        arm.preroll()
        for _ in range(3):
            arm.step(PREROLL_CHUNK)  # the same ~0.5 s under load as before the main measurement
        arm.step(args.steps)
        ms_w = arm.timed_device(1, args.steps)
        es, _, _ = arm.e2e(args.steps, args.warmup)
        weak = {"worlds_per_gpu": per_gpu, "total_worlds": per_gpu * world_size, "scaling": "weak",
                "value": DYNAMIC_BODIES * per_gpu * world_size * args.steps / (ms_w * 1e-3), "ms_per_step": ms_w / args.steps,
                "e2e_value": DYNAMIC_BODIES * per_gpu * world_size * args.steps / es, "e2e_ms_per_step": 1e3 * es / args.steps}

    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        from oracle import b2o
        threads = b2o.hardware_threads() or os.cpu_count() or 1
        r = cpu_arm(threads, 3, 1, from_t0_steps=0 if args.no_t0 else 1000)
        cpu_baseline = {"value": r["value"], "unit": "body-steps/s", "cores": threads, "kind": "port",
                        "sample": "%d Pyramid worlds x %d steps x 3 repeats, one world per host thread, persistent pool (C++ "
                                  "oracle restating box2d-rs; %.1f s of CPU work)" % (r["n_sample"], r["inner"], sum(r["secs"]) * threads),
                        "run_from_t0": r["from_t0"]}
    single_world = None
    if rank == 0 and world_size == 1 and not args.no_single_world:
        arm.close()
        single_world = single_world_leg(ctx, full=args.single_world_full)
    if rank == 0:
        line = {
            "metric": "body-steps/s, batched Pyramid worlds", "value": value, "unit": "body-steps/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(total, world_size, {
                "worlds_per_gpu": end - first,
                "l2": "inputs exceed L2: per-step state of this rank's shard is %.0f MB" % (alg_bytes_step / 1e6),
                "contacts_per_world": float(st["contacts"].mean()), "touching_per_world": float(st["touching"].mean()),
                "parallelism": "worlds sharded by contiguous index ranges, no data-path collective"}),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "validation": validation, "run_from_t0": run_t0, "weak_4096_per_gpu": weak,
            "single_world": single_world,
        }
        OUT.emit(json.dumps(line))
    arm.close()
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
    ap.add_argument("--worlds", type=int, default=4096, help="worlds in TOTAL, sharded over the ranks")
    ap.add_argument("--max-contacts", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-t0", action="store_true", help="skip the 1000-steps-from-t=0 leg")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the 4096-worlds-per-GPU comparison leg")
    ap.add_argument("--no-single-world", action="store_true", help="skip the single-large-world leg (ms/step of one world)")
    ap.add_argument("--single-world-full", action="store_true", help="longer slices of the SURVEY 8d windows (minutes)")
    ap.add_argument("--solver", default=None, help="diagnostic: lane | generic | levels ... (default: best measured)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # the CPU arm needs only the oracle: libb2gpu.so is neither built nor loaded in this process
        if rank == 0:
            try:
                subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
            except Exception as e:
                if not os.path.exists(os.path.join(ROOT, "oracle", "libb2o.so")):
                    raise
                print("bench.py: oracle rebuild skipped (%s)" % e, file=sys.stderr)
    else:
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
            run_reference(args, rank, world_size)
        else:
            run_ours(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
