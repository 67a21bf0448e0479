"""B2body's force / impulse API between steps (src/b2_body.rs:869-972, set_awake :783-801): apply_force,
apply_torque, apply_linear_impulse, apply_linear_impulse_to_center, apply_angular_impulse, set_awake — the action
interface of an RL loop.  Device world (host simulator here, CUDA in the gpu case) against the oracle, bit for bit:
the body record right after each call and the whole state after stepping; static / kinematic bodies ignore the calls,
a sleeping body accumulates nothing unless `wake`."""
import numpy as np
import pytest

import parity
from conftest import HOSTSIM_SO


def _case(ctx):
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o
    wo = b2o.B2world((0.0, -10.0))
    scenes.mixed(wo, n=60, width=12.0)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.mixed(wg, n=60, width=12.0)
    nb = wo.get_body_count()
    rng = np.random.default_rng(77)
    f = lambda lo, hi: float(np.float32(rng.uniform(lo, hi)))  # noqa: E731
    calls = 0
    for step in range(260):
        for _ in range(int(rng.integers(0, 4))):
            b = int(rng.integers(0, nb))  # body 0 is the static container: calls must be ignored
            kind = int(rng.integers(0, 9))
            vec, pt, sc, wake = (f(-60, 60), f(-60, 60)), (f(-6, 6), f(0, 12)), f(-30, 30), bool(rng.integers(0, 2))
            for w in (wo, wg):
                bd = w.body(b)
                if kind == 0: bd.apply_force(vec, pt, wake)
                elif kind == 1: bd.apply_torque(sc, wake)
                elif kind == 2: bd.apply_linear_impulse((vec[0] * 0.05, vec[1] * 0.05), pt, wake)
                elif kind == 3: bd.apply_linear_impulse_to_center((vec[0] * 0.05, vec[1] * 0.05), wake)
                elif kind == 4: bd.apply_angular_impulse(sc * 0.02, wake)
                elif kind == 5: bd.set_awake(wake)
                elif kind == 6: bd.set_damping(abs(sc) * 0.02, abs(vec[0]) * 0.01)
                elif kind == 7: bd.set_gravity_scale(sc * 0.05)
                else: bd.set_sleeping_allowed(wake)
            ro, rg = wo.body(b)._rec(), wg.body(b)._rec()
            assert ro.tobytes() == rg.tobytes(), (step, b, kind)
            calls += 1
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if step % 20 == 19:
            assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == [], step
    assert calls > 200
    # a sleeping body: without `wake` nothing accumulates, with it the body wakes and takes the impulse
    for w in (wo, wg):
        bd = w.body(5)
        bd.set_awake(False)
        bd.apply_linear_impulse((3.0, 1.0), (0.0, 1.0), False)
        bd.apply_force((9.0, 9.0), (1.0, 1.0), False)
    r = wg.body(5)._rec()
    assert not (int(r["flags"]) & abi.BODY_AWAKE) and float(r["v"][0]) == 0.0 and float(r["force"][0]) == 0.0
    for w in (wo, wg):
        w.body(5).apply_angular_impulse(0.25, True)
    r = wg.body(5)._rec()
    assert (int(r["flags"]) & abi.BODY_AWAKE) and float(r["w"]) != 0.0
    assert wo.body(5)._rec().tobytes() == r.tobytes()
    # errors: body index out of range
    assert wg.L.b2gpu_body_apply_torque(wg.h, nb + 3, 1.0, 1) == abi.E_INVALID
    assert wg.L.b2gpu_body_set_awake(None, 0, 1) == abi.E_INVALID
    wg.close()


def _gravity_case(ctx):
    """B2world::set_gravity between steps (src/b2_world.rs:232-239): the next step integrates with the new vector; bodies
    that sleep stay asleep (nobody is woken), in the plain and the large-world mode."""
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o
    for large in (0, 1):
        wo = b2o.B2world((0.0, -10.0))
        scenes.mixed(wo, n=40, width=10.0)
        wg = world.B2world((0.0, -10.0), ctx=ctx)
        scenes.mixed(wg, n=40, width=10.0)
        if large:
            wg.set_large_mode(1)
        assert wg.get_gravity() == (0.0, -10.0)
        for step in range(240):
            if step in (60, 120, 180):
                g = {60: (4.0, -6.0), 120: (0.0, 12.5), 180: (0.0, -10.0)}[step]
                for w in (wo, wg):
                    w.set_gravity(g)
                assert wg.get_gravity() == g
            if large:  # mode 1 is teacher-forced: every step starts from the oracle's state (contact order within a step differs)
                wg.upload(wo.snapshot())
                wo.step(scenes.DT, 8, 3)
                wg.step(scenes.DT, 8, 3)
                if step % 40 == 39 or step in (60, 61, 120, 121):
                    assert parity.compare_large_step(wo.snapshot(), wg.snapshot(), wo.get_stats(), wg.get_stats()) == [], (large, step)
            else:
                wo.step(scenes.DT, 8, 3)
                wg.step(scenes.DT, 8, 3)
                if step % 40 == 39 or step in (60, 61, 120, 121):
                    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == [], (large, step)
        asleep = [i for i, b in enumerate(wo.snapshot().bodies) if b["type"] == abi.DYNAMIC_BODY and not (int(b["flags"]) & abi.BODY_AWAKE)]
        if asleep:  # a sleeping body is not woken by a new gravity
            for w in (wo, wg):
                w.set_gravity((0.0, 30.0))
            assert not (int(wg.body(asleep[0])._rec()["flags"]) & abi.BODY_AWAKE)
        assert wg.L.b2gpu_world_set_gravity(None, 0.0, 0.0) == abi.E_INVALID
        wg.close()


def _batch_gravity_case(ctx):
    """b2gpu_batch_set_gravity: a different gravity per world of a batch, changed mid-run, against oracle worlds."""
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    wo = b2o.B2world((0.0, -10.0))
    scenes.mixed(wo, n=30, width=8.0)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    scenes.mixed(wg, n=30, width=8.0)
    n = 35
    bt = wg.batch(n)
    rng = np.random.default_rng(3)
    picks = {0: wo.clone(), 17: wo.clone(), n - 1: wo.clone()}
    for step in range(160):
        if step in (0, 70):
            g = np.stack([rng.uniform(-3.0, 3.0, n), rng.uniform(-15.0, -5.0, n)], 1).astype(np.float32)
            bt.set_gravity(g[:10])
            bt.set_gravity(g[10:], first=10)
            for w, o in picks.items():
                o.set_gravity((float(g[w][0]), float(g[w][1])))
        bt.step(scenes.DT, 8, 3)
        for o in picks.values():
            o.step(scenes.DT, 8, 3)
        if step % 20 == 19:
            for w, o in picks.items():
                assert parity.compare_snapshots(o.snapshot(), bt.download_world(w)) == [], (w, step)
    bt.close()
    wg.close()


def test_set_gravity_host_simulator(built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    _gravity_case(ctx)
    _batch_gravity_case(ctx)
    ctx.close()


@pytest.mark.gpu
def test_set_gravity_gpu(built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0)
    try:
        _gravity_case(ctx)
        _batch_gravity_case(ctx)
    finally:
        ctx.close()


def test_body_api_host_simulator(built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    _case(ctx)
    ctx.close()


@pytest.mark.gpu
def test_body_api_gpu(built):
    from box2d_rs_b200 import batch
    ctx = batch.Context(0)
    try:
        _case(ctx)
    finally:
        ctx.close()
