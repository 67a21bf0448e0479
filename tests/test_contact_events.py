"""begin_contact / end_contact events of a device-resident step (b2gpu_contact_events).

The oracle records the events at the points where the reference fires them (b2_contact.rs(private):201-211,
b2_contact_manager.rs(private):24-49); the library derives them from the snapshots before and after a step.  Both
lists must be equal — types, fixtures, child indices and ORDER — for every step of every scene family, with the
destroyed count given and inferred, from oracle snapshots and from a stepped world (host simulator / GPU)."""
import ctypes as C

import numpy as np
import pytest

from conftest import HOSTSIM_SO, SCENES


def _table(ev):
    return np.stack([ev["type"], ev["fixture_a"], ev["index_a"], ev["fixture_b"], ev["index_b"]], axis=1) if len(ev) else np.zeros((0, 5), np.int32)


@pytest.mark.parametrize("name", list(SCENES))
def test_events_from_oracle_snapshots(name, built):
    from box2d_rs_b200 import lib, scenes, world
    from oracle import b2o
    L = lib.load()
    recipe, gravity, steps = SCENES[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    before = wo.snapshot()
    begins = ends = 0
    for i in range(min(steps, 160)):
        wo.step(scenes.DT, 8, 3)
        after = wo.snapshot()
        ref = wo.contact_events()
        destroyed = int(wo.get_stats()["destroyed"])
        for d in (destroyed, -1):
            got = _table(world.contact_events(L, before, after, d))
            assert np.array_equal(got, ref), "step %d (destroyed=%d): %s vs %s" % (i, d, got[:4].tolist(), ref[:4].tolist())
        begins += int((ref[:, 0] == 1).sum())
        ends += int((ref[:, 0] == 2).sum())
        before = after
    assert begins > 0
    if name in ("mixed300", "pile400", "variety", "sensors", "addpair2000"):
        assert ends > 0, "the scene should separate some contacts"


def _stepped_world_case(ctx, name, steps):
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    recipe, gravity, _ = SCENES[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    total = 0
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        got = _table(wg.step_with_events(scenes.DT, 8, 3))
        assert np.array_equal(got, wo.contact_events()), i
        total += len(got)
    assert total > 0
    wg.close()


@pytest.mark.parametrize("name", ["hello_world", "sensors", "mixed300"])
def test_step_with_events_host_simulator(name, built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    _stepped_world_case(ctx, name, 90)
    ctx.close()


@pytest.mark.gpu
def test_step_with_events_gpu(built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0)
    try:
        _stepped_world_case(ctx, "sensors", 120)
        _stepped_world_case(ctx, "mixed300", 90)
    finally:
        ctx.close()


def test_truncation_and_errors(built):
    from box2d_rs_b200 import abi, lib, scenes, world
    from oracle import b2o
    L = lib.load()
    wo = b2o.B2world((0.0, -10.0))
    scenes.pyramid(wo)
    before = wo.snapshot()
    n = 0
    while n < 3:  # the step in which the rows land
        before = wo.snapshot()
        wo.step(scenes.DT, 8, 3)
        n = len(wo.contact_events())
    after = wo.snapshot()
    cb, ca = before.as_c(), after.as_c()
    out = np.zeros(2, abi.CONTACT_EVENT_DTYPE)
    assert L.b2gpu_contact_events(C.byref(cb), C.byref(ca), -1, out.ctypes.data, 2) == n  # true count, first two kept
    assert np.array_equal(_table(out), wo.contact_events()[:2])
    assert L.b2gpu_contact_events(None, C.byref(ca), -1, out.ctypes.data, 2) == abi.E_INVALID
    assert L.b2gpu_contact_events(C.byref(cb), C.byref(ca), -1, None, 2) == abi.E_INVALID
    assert L.b2gpu_contact_events(C.byref(cb), C.byref(ca), 10 ** 6, out.ctypes.data, 2) == abi.E_INVALID
    # identical snapshots: no events
    assert len(world.contact_events(L, after, after)) == 0


def test_batch_step_with_events(built):
    """Events of selected worlds of a batch (one of them perturbed) == the oracle's listener log of that world."""
    from box2d_rs_b200 import batch, scenes, world
    from oracle import b2o
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    wo = b2o.B2world((0.0, -10.0))
    scenes.pyramid(wo)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.pyramid(wg)
    bt = wg.batch(6, lane_block=4, max_contacts=800)
    o2 = wo.clone()
    o2.body(211).set_transform((2.0, 26.0), 0.4)
    bt.upload_world(5, o2.snapshot())
    total = {0: 0, 5: 0}
    for i in range(70):
        wo.step(scenes.DT, 8, 3)
        o2.step(scenes.DT, 8, 3)
        got = bt.step_with_events(scenes.DT, 8, 3, [0, 5])
        for w, o in ((0, wo), (5, o2)):
            assert np.array_equal(_table(got[w]), o.contact_events()), (i, w)
            total[w] += len(got[w])
    assert total[0] > 300 and total[5] != total[0]
    bt.close()
    wg.close()
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------------
# post_solve reports (B2island::report): fixtures, point count and impulses of every island contact, in call order
# ---------------------------------------------------------------------------------------------------------------------
def _post_solve(name, ctx, batch_mode, n=3, lane_block=1):
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    from conftest import JOINT_SCENES
    recipe, gravity, steps = {**SCENES, **JOINT_SCENES}[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    bt = wg.batch(n, lane_block=lane_block) if batch_mode else None
    seen = 0
    for i in range(min(steps, 150)):
        wo.step(scenes.DT, 8, 3)
        if bt is not None:
            bt.step(scenes.DT, 8, 3)
        else:
            wg.step(scenes.DT, 8, 3)
        if i % 7 == 0 or i < 3:
            ref = wo.post_solve_events()
            got = bt.post_solve_events(n - 1) if bt is not None else wg.post_solve_events()
            assert len(ref) == len(got), "step %d: %d vs %d reports" % (i, len(ref), len(got))
            for f in ("fixture_a", "index_a", "fixture_b", "index_b", "count"):
                assert np.array_equal(ref[f], got[f]), "step %d: %s" % (i, f)
            for f in ("normal_impulses", "tangent_impulses"):
                assert np.array_equal(ref[f].view(np.uint32), got[f].view(np.uint32)), "step %d: %s" % (i, f)
            seen += len(ref)
    assert seen > 0
    if bt is not None:
        bt.close()
    wg.close()


@pytest.mark.parametrize("name,batch_mode", [("pyramid", True), ("mixed300", False), ("variety", True), ("joints_mix", False)])
def test_post_solve_reports_match_the_oracle(name, batch_mode, built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    _post_solve(name, ctx, batch_mode)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,batch_mode", [("pyramid", True), ("variety", False), ("joints_mix", True)])
def test_post_solve_reports_match_the_oracle_gpu(name, batch_mode, built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0)
    _post_solve(name, ctx, batch_mode, n=34, lane_block=0)  # 32-world memory blocks: the shared-memory kernels
    ctx.close()
