"""World queries: B2world::ray_cast with the closest-hit callback and B2world::query_aabb,
device path (host simulator here, CUDA in the gpu-marked cases) against the CPU restatement
(oracle/b2o_query.hpp), bit for bit — fixture, child, fraction, point and normal of every ray, and the proxies
of every box in the reference's report order (as a set in large-world mode 1, whose walk is the LBVH)."""
import numpy as np
import pytest

import parity  # noqa: F401
from conftest import HOSTSIM_SO, SCENES


def make_queries(seed, n, lo, hi):
    rng = np.random.default_rng(seed)
    p1 = rng.uniform(lo, hi, (n, 2))
    ang = rng.uniform(0, 2 * np.pi, n)
    length = rng.uniform(0.2, 0.6 * (hi[0] - lo[0]), n)
    rays = np.concatenate([p1, p1 + np.stack([np.cos(ang), np.sin(ang)], 1) * length[:, None]], 1).astype(np.float32)
    idx = np.arange(n)
    hor, ver, short = idx % 7 == 0, (idx % 11 == 3) & (idx % 7 != 0), (idx % 13 == 5) & (idx % 7 != 0) & (idx % 11 != 3)
    rays[hor, 3] = rays[hor, 1]          # horizontal
    rays[ver, 2] = rays[ver, 0]          # vertical
    rays[short, 2:] = rays[short, :2] + np.float32(0.01)  # very short
    c = rng.uniform(lo, hi, (n, 2))
    h = rng.uniform(0.05, 3.0, (n, 2))
    boxes = np.concatenate([c - h, c + h], 1).astype(np.float32)
    return rays, boxes


def check_queries(wo, wg, seed, lo, hi, ordered=True, n=400):
    rays, boxes = make_queries(seed, n, np.array(lo), np.array(hi))
    ref, got = wo.ray_cast_closest(rays), wg.ray_cast_closest(rays)
    assert np.array_equal(ref["fixture"], got["fixture"]) and np.array_equal(ref["child_index"], got["child_index"])
    for f in ("fraction", "point", "normal"):
        assert np.array_equal(ref[f].view(np.uint32), got[f].view(np.uint32)), f
    (rh, rc), (gh, gc) = wo.query_aabb(boxes, 256), wg.query_aabb(boxes, 256)
    assert np.array_equal(rc, gc)
    if ordered:
        assert rh == gh
    else:
        assert [sorted(a) for a in rh] == [sorted(b) for b in gh]
    return int((ref["fixture"] >= 0).sum()), int(rc.sum())


BOUNDS = {"pyramid": ((-15, -1), (15, 28)), "mixed300": ((-16, -1), (16, 40)), "variety": ((-14, -2), (14, 14)),
          "sensors": ((-13, -1), (13, 12)), "terrain": ((-120, -6), (120, 20)), "addpair2000": ((-110, -12), (5, 22))}


def run_scene(name, ctx, mode, steps_between=40, rounds=3):
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    recipe, gravity, _ = SCENES[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    if mode:
        wg.set_large_mode(mode)
    hits = boxes = 0
    lo, hi = BOUNDS[name]
    for r in range(rounds):  # as built, then after some stepping
        if mode == 1 and r > 0:
            wg.upload(wo.snapshot())  # mode 1 diverges free-running: query the same state
            wg.set_large_mode(1)
        a, b = check_queries(wo, wg, 100 * r + len(name), lo, hi, ordered=mode != 1)
        hits += a
        boxes += b
        for _ in range(steps_between):
            wo.step(scenes.DT, 8, 3)
            wg.step(scenes.DT, 8, 3)
    assert hits > 50 and boxes > 200, (hits, boxes)  # the queries really hit things
    wg.close()


@pytest.fixture(scope="module")
def hctx(built):
    from box2d_rs_b200 import batch
    c = batch.Context(0, lib_path=HOSTSIM_SO)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(BOUNDS))
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_queries_host_simulator(name, mode, hctx):
    run_scene(name, hctx, mode)


def test_ray_cast_known_answers(hctx):
    """Hand-computable cases through the oracle and the device path: a circle hit head-on, the polygon entry
    threshold 0.032 of the Rust port (b2_polygon_shape.rs(private):241), a one-sided edge from its back side."""
    from box2d_rs_b200 import abi, world
    from oracle import b2o

    def build(w):
        g = w.create_body(abi.BodyDef())
        g.create_fixture_by_shape(w.shapes.circle(1.0, (5.0, 0.0)), 0.0)                       # fixture 0
        g.create_fixture_by_shape(w.shapes.polygon_box(1.0, 1.0, (0.0, 10.0), 0.0), 0.0)        # fixture 1
        g.create_fixture_by_shape(w.shapes.edge_one_sided((-3.0, 20.0), (-1.0, 20.0), (1.0, 20.0), (3.0, 20.0)), 0.0)  # fixture 2

    wo = b2o.B2world((0.0, -10.0))
    build(wo)
    wg = world.B2world((0.0, -10.0), ctx=hctx)
    build(wg)
    rays = np.array([[0.0, 0.0, 10.0, 0.0],      # circle: enters at x = 4 -> fraction 0.4, normal (-1, 0)
                     [-5.0, 10.0, 5.0, 10.0],    # box: enters at x = -1 -> fraction 0.4, normal (-1, 0)
                     [-1.02, 10.0, 5.0, 10.0],   # box entered at fraction 0.0033 < 0.032: the port reports no hit
                     [0.0, 25.0, 0.0, 15.0],     # one-sided edge from the side its normal faces
                     [0.0, 15.0, 0.0, 25.0]],    # ... and from behind: no hit
                    np.float32)
    ref, got = wo.ray_cast_closest(rays), wg.ray_cast_closest(rays)
    for f in ("fixture", "child_index"):
        assert np.array_equal(ref[f], got[f])
    for f in ("fraction", "point", "normal"):
        assert np.array_equal(ref[f].view(np.uint32), got[f].view(np.uint32))
    assert ref["fixture"].tolist()[:3] == [0, 1, -1]
    assert abs(float(ref["fraction"][0]) - 0.4) < 1e-6 and np.allclose(ref["normal"][0], (-1.0, 0.0))
    assert abs(float(ref["fraction"][1]) - 0.4) < 1e-6 and np.allclose(ref["normal"][1], (-1.0, 0.0))
    assert sorted(ref["fixture"].tolist()[3:]) == [-1, 2]
    with pytest.raises(Exception):
        wg.ray_cast_closest(np.array([[1.0, 1.0, 1.0, 1.0]], np.float32))  # the reference asserts p1 != p2
    wg.close()


def test_batch_ray_cast(hctx):
    """b2gpu_batch_ray_cast_closest: the same rays in every replica of a batch, one perturbed world differs."""
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    wo = b2o.B2world((0.0, -10.0))
    scenes.pyramid(wo)
    wg = world.B2world((0.0, -10.0), ctx=hctx)
    scenes.pyramid(wg)
    bt = wg.batch(5, lane_block=4, max_contacts=800)
    o2 = wo.clone()
    o2.body(211).set_transform((2.0, 26.0), 0.4)
    bt.upload_world(3, o2.snapshot())
    for _ in range(30):
        bt.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
        o2.step(scenes.DT, 8, 3)
    rays, _ = make_queries(5, 200, np.array((-15, -1)), np.array((15, 28)))
    got = bt.ray_cast_closest(np.broadcast_to(rays, (5,) + rays.shape))
    for w, o in ((0, wo), (3, o2), (4, wo)):
        ref = o.ray_cast_closest(rays)
        assert np.array_equal(ref["fixture"], got[w]["fixture"])
        assert np.array_equal(ref["fraction"].view(np.uint32), got[w]["fraction"].view(np.uint32))
        assert np.array_equal(ref["normal"].view(np.uint32), got[w]["normal"].view(np.uint32))
    bt.close()
    wg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pyramid", "mixed300", "variety", "terrain", "addpair2000"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_queries_gpu(name, mode, built):
    from box2d_rs_b200 import batch
    c = batch.Context(0)
    try:
        run_scene(name, c, mode)
    finally:
        c.close()


def _batch_query_aabb_case(ctx, lane_block):
    """b2gpu_batch_query_aabb: different boxes in every world of a batch (one world perturbed), each world's
    answer == the oracle's query_aabb of that world: counts, and (fixture, child) in the reference's report order;
    truncation at max_hits keeps the true count."""
    from box2d_rs_b200 import abi, lib, scenes, world
    from oracle import b2o
    wo = b2o.B2world((0.0, -10.0))
    scenes.pyramid(wo)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.pyramid(wg)
    n_worlds, per_world = 6, 60
    bt = wg.batch(n_worlds, lane_block=lane_block, max_contacts=800)
    o2 = wo.clone()
    o2.body(211).set_transform((2.0, 26.0), 0.4)
    bt.upload_world(4, o2.snapshot())
    for _ in range(30):
        bt.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
        o2.step(scenes.DT, 8, 3)
    boxes = np.stack([make_queries(50 + w, per_world, np.array((-15, -1)), np.array((15, 28)))[1] for w in range(n_worlds)])
    hits, counts = bt.query_aabb(boxes, max_hits=128)
    total = 0
    for w in range(n_worlds):
        ref_hits, ref_counts = (o2 if w == 4 else wo).query_aabb(boxes[w], 128)
        assert np.array_equal(ref_counts, counts[w]), w
        for i in range(per_world):
            assert [tuple(int(v) for v in h) for h in hits[w, i, :counts[w, i]]] == ref_hits[i], (w, i)
        total += int(ref_counts.sum())
    assert total > 500 and counts.max() > 6
    small_hits, small_counts = bt.query_aabb(boxes, max_hits=3)
    assert np.array_equal(small_counts, counts)
    keep = np.arange(3)[None, None, :] < np.minimum(counts, 3)[:, :, None]  # entries past a box's count are unspecified
    assert np.array_equal(small_hits[keep], hits[:, :, :3][keep])
    # no boxes: nothing to do; NULL batch: error code
    assert bt.L.b2gpu_batch_query_aabb(bt.h, boxes.ctypes.data, 0, 4, counts.ctypes.data, hits.ctypes.data) == 0
    assert bt.L.b2gpu_batch_query_aabb(None, boxes.ctypes.data, 1, 4, counts.ctypes.data, hits.ctypes.data) == abi.E_INVALID
    # the one-world call refuses a batch and says where to go
    with pytest.raises(lib.B2gpuError):
        lib.check(bt.L, bt.L.b2gpu_batch_query_aabb(bt.h, boxes.ctypes.data, -1, 4, counts.ctypes.data, hits.ctypes.data))
    bt.close()
    wg.close()


@pytest.mark.parametrize("lane_block", [1, 4, 32])
def test_batch_query_aabb(lane_block, hctx):
    _batch_query_aabb_case(hctx, lane_block)


@pytest.mark.gpu
def test_batch_query_aabb_gpu(built):
    from box2d_rs_b200 import batch
    c = batch.Context(0)
    try:
        _batch_query_aabb_case(c, 32)
        _batch_query_aabb_case(c, 1)
    finally:
        c.close()
