#!/usr/bin/env python
"""Development aid: where a jointed batch spends its step (per-stage events), and what one call costs when it carries 1 or
many steps, with and without stream groups.  Not a benchmark."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from box2d_rs_b200 import scenes, world
from box2d_rs_b200.batch import Context
stream = torch.cuda.Stream()
ctx = Context(0, stream=stream.cuda_stream)
for name, rec in (("car", lambda w: scenes.car(w)), ("joints_mix", scenes.joints_mix), ("pyramid", scenes.pyramid)):
    for solver in (None, "one_stream"):
        wg = world.B2world((0.0, -10.0), ctx=ctx)
        rec(wg)
        bt = wg.batch(4096, solver=solver)
        for _ in range(60): bt.step(scenes.DT, 8, 3)
        ctx.sync()
        out = {}
        for label, calls, per in (("200x1", 200, 1), ("10x20", 10, 20), ("1x200", 1, 200)):
            bt.step(scenes.DT, 8, 3, per)  # build the graph of this signature
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(calls): bt.step(scenes.DT, 8, 3, per)
                e1.record(stream)
            torch.cuda.synchronize()
            out[label] = (round(e0.elapsed_time(e1) / (calls * per), 4), round(1e3 * (time.perf_counter() - t0) / (calls * per), 4))
        ctx.set_profiling(True)
        for _ in range(20): bt.step(scenes.DT, 8, 3)
        ctx.sync()
        ms = {k: round(v[0] / 20, 4) for k, v in ctx.stage_times().items() if v[0] / 20 > 0.02}
        ctx.set_profiling(False)
        print(name, solver, "ms/step (device, wall) by calls x steps:", out, "| stages sum", round(sum(ms.values()), 3), ms)
        bt.close(); wg.close()
