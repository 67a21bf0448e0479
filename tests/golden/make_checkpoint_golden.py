"""Regenerates tests/golden/hello_world_step30.b2snap: the oracle's hello_world scene after 30 steps, written by
b2gpu_snapshot_save.  The committed file pins the on-disk format (version 1): a later build must still read it
(tests/test_checkpoint.py::test_committed_file_still_loads), and its content must agree with hello_world.npz.
    python tests/golden/make_checkpoint_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from box2d_rs_b200 import checkpoint, scenes
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    scenes.hello_world(w)
    for _ in range(30):
        w.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS)
    path = os.path.join(HERE, "hello_world_step30.b2snap")
    checkpoint.save(w.snapshot(), path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
