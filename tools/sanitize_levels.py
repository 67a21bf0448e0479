import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from box2d_rs_b200 import scenes, world
from box2d_rs_b200.batch import Context
ctx = Context(0)
for name, fn, steps in (("pile", lambda w: scenes.pile(w, n=300, width=10.0), 60), ("pyramid", scenes.pyramid, 50)):
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    fn(wg)
    wg.set_large_mode(1)
    wg.set_level_threshold(4)
    lv = 0
    for i in range(steps):
        wg.step(scenes.DT, 8, 3)
        lv = max(lv, int(wg.get_stats()["solver_levels"]))
    print(name, "levels", lv, "status", int(wg.get_stats()["status"]))
    wg.close()
