"""Regenerates tests/golden/*.npz with the CPU oracle (oracle/libb2o.so).

The reference (box2d-rs) ships no golden vectors and cannot be built in this image (no Rust
toolchain), so these fixtures pin the *oracle's* outputs: body state (c.x c.y a v.x v.y w xf.p) and
the contact set (fixture_a, index_a, fixture_b, index_b, flags, manifold point count and feature
ids) at chosen steps of the BASELINE.json scenes at small sizes.  They guard the oracle, the host
simulator and the CUDA path against silent drift:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (scene key in conftest.SCENES, steps at which state is recorded)
    "pyramid": ("pyramid", (1, 10, 60, 200, 300)),
    "hello_world": ("hello_world", (1, 30, 60)),
    "mixed300": ("mixed300", (1, 50, 150, 300)),
    "pile400": ("pile400", (1, 50, 150)),
    "addpair2000": ("addpair2000", (1, 40, 120)),
    "variety": ("variety", (1, 60, 200, 400)),
    "sensors": ("sensors", (1, 60, 150, 300)),
    "terrain": ("terrain", (1, 80, 160, 260)),
}


def contact_table(snap):
    c = snap.contacts
    m = c["manifold"]
    return np.stack([c["fixture_a"], c["index_a"], c["fixture_b"], c["index_b"], c["flags"].astype(np.int64),
                     m["point_count"], m["type"], m["points"]["id"][:, 0].astype(np.int64),
                     m["points"]["id"][:, 1].astype(np.int64)], axis=1).astype(np.int64)


def record(world, steps, step_fn):
    out = {}
    done = 0
    for s in steps:
        while done < s:
            step_fn()
            done += 1
        out["state_%d" % s] = world.body_state() if hasattr(world, "body_state") else None
        out["contacts_%d" % s] = contact_table(world.snapshot())
    return out


def main():
    from conftest import SCENES
    from box2d_rs_b200 import scenes
    from oracle import b2o
    only = sys.argv[1:]  # optional: regenerate only the named cases (a new scene must not rewrite the old pins)
    for name, (key, steps) in CASES.items():
        if only and name not in only:
            continue
        recipe, gravity, _ = SCENES[key]
        w = b2o.B2world(gravity)
        recipe(scenes, w)
        data = record(w, steps, lambda: w.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), steps=np.array(steps), **data)
        print(name, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
