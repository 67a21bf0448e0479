#!/usr/bin/env python
"""Small driver for sanitizer / profiler runs of the large-world mode: builds a scene, steps it, prints stats.

  compute-sanitizer --tool memcheck python tools/large_smoke.py --scene pile --n 400 --steps 40 [--mode 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="pile", choices=["pile", "mixed", "addpair", "terrain", "variety"])
    ap.add_argument("--n", type=int, default=400)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--mode", type=int, default=1, help="1 large-world mode, 2 exact-order large-world mode")
    args = ap.parse_args()
    from box2d_rs_b200 import scenes, world
    gravity = (0.0, 0.0) if args.scene == "addpair" else (0.0, -10.0)
    w = world.B2world(gravity)
    if args.scene == "pile":
        scenes.pile(w, n=args.n, width=max(12.0, 0.004 * args.n))
    elif args.scene == "mixed":
        scenes.mixed(w, n=args.n, width=max(30.0, 0.01 * args.n))
    elif args.scene == "addpair":
        scenes.add_pair(w, n=args.n)
    elif args.scene == "terrain":
        scenes.terrain(w)
    else:
        scenes.variety(w)
    w.set_large_mode(args.mode)
    for _ in range(args.steps):
        w.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
    st = w.get_stats()
    print({k: int(st[k]) for k in ("status", "contacts", "touching", "islands", "awake_bodies")})
    w.close()


if __name__ == "__main__":
    main()
