"""CPU tests (no GPU) of the large-world mode (b2g_large.h) on the test-only host simulator: every step is
the oracle's step of the same state — teacher-forced: upload the oracle's state, step
both once, compare everything bit for bit, the contacts created in that step as a set."""
import numpy as np
import pytest

import parity
from conftest import HOSTSIM_SO, SCENES


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


def teacher_forced(wo, bt, steps, every, world_index=0):
    """Returns the largest b2gpu_step_stats.solver_levels seen (0: no island took the level-scheduled form)."""
    from box2d_rs_b200 import scenes
    levels = 0
    for i in range(steps):
        if i % every == 0:
            bt.upload_world(world_index, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            st = bt.stats()[world_index]
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(world_index), wo.get_stats(), st)
            assert bad == [], "step %d: %s" % (i, bad[:6])
            levels = max(levels, int(st["solver_levels"]))
        else:
            wo.step(scenes.DT, 8, 3)
    return levels


@pytest.mark.parametrize("name,every", [("hello_world", 1), ("pyramid", 2), ("mixed300", 2), ("pile400", 2), ("variety", 1),
                                        ("sensors", 1), ("addpair2000", 3), ("terrain", 2)])
def test_large_mode_teacher_forced(name, every, ctx):
    wo, wg, steps = _pair(name, ctx)
    bt = wg.batch(1, lane_block=1, solver='large')
    teacher_forced(wo, bt, steps, every)
    bt.close()
    wg.close()


def test_large_mode_wake_chain(ctx):
    """A sleeping stack hit by a fast body: the ordered wake fix-up after the flat narrowphase."""
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o

    def build(w):
        ground = w.create_body(abi.BodyDef())
        ground.create_fixture_by_shape(w.shapes.edge_two_sided((-20.0, 0.0), (20.0, 0.0)), 0.0)
        box = w.shapes.polygon_box(0.5, 0.5)
        for i in range(6):
            b = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.51 + 1.02 * i)))
            b.create_fixture_by_shape(box, 1.0)
        bullet = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(-12.0, 3.0), allow_sleep=0))
        bullet.create_fixture_by_shape(w.shapes.circle(0.4), 2.0)
        return bullet

    wo = b2o.B2world((0.0, -10.0))
    bo = build(wo)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    build(wg)
    bt = wg.batch(1, lane_block=1, solver='large')
    for i in range(260):
        if i == 200:
            assert int(wo.get_stats()["awake_bodies"]) <= 1
            bo.set_transform((-6.0, 3.0), 0.0)
            bo.set_linear_velocity((25.0, 0.0))
        if i < 5 or i >= 195:
            bt.upload_world(0, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(0), wo.get_stats(), bt.stats()[0])
            assert bad == [], "step %d: %s" % (i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    bt.close()
    wg.close()


def test_large_mode_world_api_and_tree_refit(ctx):
    """b2gpu_world_set_large_mode: free-running is deterministic, keeps the contact SET of a valid Box2D run
    consistent (every live contact's fat boxes overlap, no duplicates), and the downloaded snapshot carries a
    valid bounding hierarchy so that bodies created afterwards find their pairs."""
    from box2d_rs_b200 import abi, scenes, world
    runs = []
    for _ in range(2):
        wg = world.B2world((0.0, -10.0), ctx=ctx)
        scenes.pile(wg, n=300, width=10.0)
        wg.set_large_mode(True)
        for _ in range(60):
            wg.step(scenes.DT, 8, 3)
        runs.append(wg)
    a, b = runs[0].snapshot(), runs[1].snapshot()
    assert parity.compare_snapshots(a, b) == []
    nodes = a.nodes
    for i in np.nonzero(nodes["height"] > 0)[0]:
        c1, c2 = nodes["child1"][i], nodes["child2"][i]
        lo = np.minimum(nodes["aabb"][c1][:2], nodes["aabb"][c2][:2])
        hi = np.maximum(nodes["aabb"][c1][2:], nodes["aabb"][c2][2:])
        assert np.array_equal(nodes["aabb"][i][:2], lo) and np.array_equal(nodes["aabb"][i][2:], hi)
    keys = set()
    for c in a.contacts:
        k = (int(c["fixture_a"]), int(c["index_a"]), int(c["fixture_b"]), int(c["index_b"]))
        assert k not in keys and (k[2], k[3], k[0], k[1]) not in keys
        keys.add(k)
    # a body dropped in after the download: its proxy is inserted on the host into the refitted tree
    wg = runs[0]
    nb = wg.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.5)))
    nb.create_fixture_by_shape(wg.shapes.circle(0.3), 1.0)
    before = wg.get_contact_count()
    wg.step(scenes.DT, 8, 3)
    assert wg.get_contact_count() > before
    assert int(wg.get_stats()["status"]) == 0
    for w in runs:
        w.close()


def test_large_mode_wake_cascade_through_touching_changes(ctx):
    """Sleeping boxes teleported apart while asleep (set_transform does not wake): when a ball then touches the
    first one, every stale TOUCHING flag down the former stack flips in ONE collide call, each flip waking the
    next body — the cascade the reference resolves by its newest-first loop order (LwWakeK's fixpoint rounds)."""
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o

    def build(w):
        ground = w.create_body(abi.BodyDef())
        ground.create_fixture_by_shape(w.shapes.edge_two_sided((-30.0, 0.0), (30.0, 0.0)), 0.0)
        box = w.shapes.polygon_box(0.5, 0.5)
        for i in range(7):
            b = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.51 + 1.02 * i)))
            b.create_fixture_by_shape(box, 1.0)
        ball = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(-15.0, 3.0), allow_sleep=0, gravity_scale=0.0))
        ball.create_fixture_by_shape(w.shapes.circle(0.4), 2.0)

    wo = b2o.B2world((0.0, -10.0))
    build(wo)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    build(wg)
    bt = wg.batch(1, lane_block=1, solver='large')
    woke_in_one_step = 0
    for i in range(240):
        if i == 200:
            assert int(wo.get_stats()["awake_bodies"]) <= 1
            # lift the sleeping boxes 1..6 slightly apart (their fat boxes still overlap, so the contacts survive and
            # keep a stale TOUCHING flag); the ball is put right on top of the uppermost one
            for k in range(1, 7):
                wo.body(1 + k).set_transform((0.0, 0.51 + 1.02 * k + 0.03 * k), 0.0)
            wo.body(8).set_transform((0.0, 0.51 + 1.02 * 6 + 0.18 + 0.5 + 0.39), 0.0)
        if i >= 198:
            before = int(wo.get_stats()["awake_bodies"])
            bt.upload_world(0, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(0), wo.get_stats(), bt.stats()[0])
            assert bad == [], "step %d: %s" % (i, bad[:6])
            woke_in_one_step = max(woke_in_one_step, int(wo.get_stats()["awake_bodies"]) - before)
        else:
            wo.step(scenes.DT, 8, 3)
    assert woke_in_one_step >= 4, woke_in_one_step  # the cascade really happened
    bt.close()
    wg.close()


@pytest.mark.parametrize("name", list(SCENES))
def test_large_mode_exact_order_free_running(name, ctx):
    """b2gpu_world_set_large_mode(w, 2): the data-parallel stages with the replica tree kept (sequential
    re-insertion, queries on that tree) create contacts in the reference's order, so the whole snapshot — tree
    included — stays bit-identical to the oracle free-running, like the exact mode."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    wg.set_large_mode(2)
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 2 or i % 40 == 39 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + \
                [b for b in parity.compare_stats(wo.get_stats(), wg.get_stats()) if "island_bodies" not in b]
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()


# ---- level-scheduled sweeps of giant islands (b2g_levels.h).  The host simulator runs a CTA functor with one thread, so the
# constraints of an island are visited in LEVEL order instead of list order: the commutation argument, tested bit for bit.
@pytest.mark.parametrize("name,every", [("pyramid", 2), ("mixed300", 2), ("pile400", 2), ("variety", 1), ("addpair2000", 3), ("terrain", 2)])
def test_level_order_teacher_forced(name, every, ctx):
    wo, wg, steps = _pair(name, ctx)
    bt = wg.batch(1, lane_block=1, solver='large')
    bt.set_level_threshold(6)
    assert teacher_forced(wo, bt, steps, every) > 0  # islands really took the level form
    bt.close()
    wg.close()


@pytest.mark.parametrize("name", ["pyramid", "pile400", "mixed300"])
def test_level_order_free_running(name, ctx):
    """Mode 2 (reference contact order) with every island of >= 4 contacts swept level by level: the whole snapshot stays
    bit-identical to the oracle free-running, through island rebuilds and cached islands alike."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    wg.set_large_mode(2)
    wg.set_level_threshold(4)
    used = 0
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 2 or i % 40 == 39 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + \
                [b for b in parity.compare_stats(wo.get_stats(), wg.get_stats()) if "island_bodies" not in b]
            assert bad == [], "step %d: %s" % (i, bad[:6])
            used = max(used, int(wg.get_stats()["solver_levels"]))
    assert used > 0
    wg.close()


def test_level_threshold_off_and_more_giants_than_slots(ctx):
    """A negative threshold switches the level form off; with more eligible islands than CTA slots (16) the surplus keeps
    the one-thread form — same bits either way."""
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o

    def build(w):
        ground = w.create_body(abi.BodyDef())
        ground.create_fixture_by_shape(w.shapes.edge_two_sided((-100.0, 0.0), (100.0, 0.0)), 0.0)
        box = w.shapes.polygon_box(0.5, 0.5)
        for s in range(24):  # 24 separate stacks of 4 boxes: 24 islands of 4 contacts once they rest
            for i in range(4):
                b = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(-60.0 + 5.0 * s, 0.51 + 1.02 * i), allow_sleep=0))
                b.create_fixture_by_shape(box, 1.0)

    for thr in (-1, 3):
        wo = b2o.B2world((0.0, -10.0))
        build(wo)
        wg = world.B2world((0.0, -10.0), ctx=ctx)
        build(wg)
        wg.set_large_mode(2)
        wg.set_level_threshold(thr)
        for i in range(40):
            wo.step(scenes.DT, 8, 3)
            wg.step(scenes.DT, 8, 3)
        assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
        st = wg.get_stats()
        assert int(st["islands"]) == 24
        assert (int(st["solver_levels"]) > 0) == (thr > 0)
        wg.close()
