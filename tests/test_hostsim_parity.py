"""CPU tests (no GPU) of the host-side logic: the device stages, compiled for the host by the test-only
simulator (tests/hostsim, every kernel launch = a plain loop), must reproduce the oracle bit for bit —
free-running, teacher-forced, batched with every memory-block size — and the committed golden fixtures."""
import os

import numpy as np
import pytest

import parity
from conftest import HOSTSIM_SO, ROOT, SCENES


@pytest.fixture(scope="module")
def ctx(built):
    from box2d_rs_b200 import batch
    c = batch.Context(0, lib_path=HOSTSIM_SO)
    yield c
    c.close()


def _pair(name, ctx):
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    recipe, gravity, steps = SCENES[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    return wo, wg, steps


@pytest.mark.parametrize("name", list(SCENES))
def test_builder_and_free_running(name, ctx):
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 2 or i % 40 == 39 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()


@pytest.mark.parametrize("lane_block", [1, 2, 8, 32])
def test_batch_memory_blocks(lane_block, ctx):
    """Blocked world-minor layout: any block size gives the same worlds (perturbed Pyramid replicas)."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("pyramid", ctx)
    n = 37
    batch = wg.batch(n, lane_block=lane_block, max_contacts=800)
    picks = [0, 5, 36]
    oracles = {}
    for w in picks:
        o = wo.clone()
        o.body(211).set_transform((3.6875 + 0.02 * w - 0.3, 24.5), 0.03 * w)
        batch.upload_world(w, o.snapshot())
        oracles[w] = o
    for _ in range(45):
        batch.step(scenes.DT, 8, 3)
        for o in oracles.values():
            o.step(scenes.DT, 8, 3)
    for w, o in oracles.items():
        assert parity.compare_snapshots(o.snapshot(), batch.download_world(w)) == []
    # untouched replicas equal the unperturbed oracle
    for _ in range(45):
        wo.step(scenes.DT, 8, 3)
    assert parity.compare_snapshots(wo.snapshot(), batch.download_world(1)) == []
    batch.close()
    wg.close()


def test_teacher_forced(ctx):
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("mixed300", ctx)
    batch = wg.batch(3, lane_block=1)
    for i in range(150):
        if i % 5 == 0:
            batch.upload_world(2, wo.snapshot())
            batch.step(scenes.DT, 8, 3)
            wo.step(scenes.DT, 8, 3)
            bad = parity.compare_snapshots(wo.snapshot(), batch.download_world(2)) + \
                parity.compare_stats(wo.get_stats(), batch.stats()[2])
            assert bad == [], "step %d: %s" % (i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    batch.close()
    wg.close()


def test_wake_up_chain_in_list_order(ctx):
    """A sleeping pile hit by a fast body: collide must wake bodies in contact-list order (the ordered
    fix-up pass after the flat narrowphase) — compare the whole episode with the oracle."""
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o

    def build(w):
        ground = w.create_body(abi.BodyDef())
        ground.create_fixture_by_shape(w.shapes.edge_two_sided((-20.0, 0.0), (20.0, 0.0)), 0.0)
        box = w.shapes.polygon_box(0.5, 0.5)
        bodies = []
        for i in range(6):
            b = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.51 + 1.02 * i)))
            b.create_fixture_by_shape(box, 1.0)
            bodies.append(b)
        bullet = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(-12.0, 3.0), allow_sleep=0))
        bullet.create_fixture_by_shape(w.shapes.circle(0.4), 2.0)
        return bullet

    wo = b2o.B2world((0.0, -10.0))
    bo = build(wo)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    bg = build(wg)
    for i in range(420):
        if i == 200:  # the stack is asleep by now; fire the ball at it
            assert int(wo.get_stats()["awake_bodies"]) <= 1
            bo.set_transform((-6.0, 3.0), 0.0)
            bg.set_transform((-6.0, 3.0), 0.0)
            bo.set_linear_velocity((25.0, 0.0))
            bg.set_linear_velocity((25.0, 0.0))
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i % 10 == 9 or 200 <= i < 230:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()


def test_forces_and_velocity_inputs(ctx):
    """b2gpu_batch_set_forces / set_linear_velocity mirror apply_force_to_center / set_linear_velocity."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("pyramid", ctx)
    batch = wg.batch(4, lane_block=4, max_contacts=800)
    forces = np.zeros((4, batch.body_count, 3), np.float32)
    forces[3, 211, 0] = 250.0
    o = wo.clone()
    v = np.zeros((4, 2), np.float32)
    v[2] = (1.5, 0.0)
    batch.set_linear_velocity(100, v)
    o2 = wo.clone()
    o2.body(100).set_linear_velocity((1.5, 0.0))
    for _ in range(40):
        batch.set_forces(forces)
        o.body(211).apply_force_to_center((250.0, 0.0), wake=False)
        batch.step(scenes.DT, 8, 3)
        o.step(scenes.DT, 8, 3)
        o2.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
    st = batch.body_state()
    assert np.array_equal(st[3].view(np.uint32), o.body_state().view(np.uint32))
    assert np.array_equal(st[2].view(np.uint32), o2.body_state().view(np.uint32))
    assert np.array_equal(st[0].view(np.uint32), wo.body_state().view(np.uint32))
    batch.close()
    wg.close()


@pytest.mark.parametrize("name", ["pyramid", "mixed300", "terrain"])
def test_simulator_matches_golden(name, ctx):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    from box2d_rs_b200 import scenes, world
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    recipe, gravity, _ = SCENES[name]
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    batch = wg.batch(1, lane_block=1)

    class View:
        def body_state(self):
            return batch.body_state()[0]

        def snapshot(self):
            return batch.download_world(0)

    got = make_golden.record(View(), [int(s) for s in g["steps"]], lambda: batch.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS))
    for k, v in got.items():
        if k.startswith("state"):
            assert np.array_equal(g[k].view(np.uint32), v.view(np.uint32)), k
        else:
            assert np.array_equal(g[k], v), k
    batch.close()
    wg.close()


@pytest.mark.parametrize("warm,block,sleep", [(False, True, True), (True, False, True), (False, False, False)])
def test_world_flags(warm, block, sleep, ctx):
    """set_warm_starting / G_BLOCK_SOLVE / set_allow_sleeping, plus a dt = 0 step (collide only) in between."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("variety", ctx)
    for w in (wo, wg):
        w.set_warm_starting(warm)
        w.set_block_solve(block)
        w.set_allow_sleeping(sleep)
    for i in range(150):
        dt = 0.0 if i in (40, 41, 90) else scenes.DT
        wo.step(dt, 8, 3)
        wg.step(dt, 8, 3)
        if i % 30 == 29 or i in (40, 41, 42):
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()
