# Step time (ms) of the 4096-world Pyramid batch in windows of 40 steps, for three perturbation sets:
# shows the falling / settling / settled phases that bench.py's pre-roll skips.
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from box2d_rs_b200 import scenes, world, sharding
from box2d_rs_b200.batch import Context
ctx = Context(0)
wg = world.B2world((0.0, -10.0), ctx=ctx); scenes.pyramid(wg); wg.set_allow_sleeping(False)
for rank in (0, 1, 5):
    b = wg.batch(4096, max_contacts=1024)
    b.set_linear_velocity(211, sharding.perturbation(rank * 4096, 4096, 0xB2D + 3))
    out = []
    for i in range(30):
        ctx.sync(); t = time.perf_counter(); b.step(scenes.DT, 8, 3, 40); ctx.sync()
        out.append(round((time.perf_counter() - t) / 40 * 1e3, 2))
    print(rank, out)
    b.close()
