"""GPU parity tests: the CUDA path, called through the C ABI (libb2gpu.so), against the CPU oracle on
the same inputs.  Bar (BASELINE.json north_star): pair/contact sets and every integer field
bit-exact, fp32 state within 1e-5 relative — the engine is FMA-free and evaluates sin/cos with the
restated libm algorithm, so the tests demand bit-equality (rtol = atol = 0) of the whole snapshot."""
import numpy as np
import pytest

import parity
from conftest import SCENES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    from box2d_rs_b200 import batch
    c = batch.Context(0)
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


def test_device_sincos_matches_libm(ctx):
    from oracle import b2o
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2**32, size=1 << 20, dtype=np.uint64).astype(np.uint32)
    a = bits.view(np.float32)
    a = a[np.isfinite(a)]
    a = np.concatenate([a, rng.uniform(-7.0, 7.0, 1 << 20).astype(np.float32),
                        rng.uniform(-1e-3, 1e-3, 1 << 16).astype(np.float32),
                        np.array([0.0, -0.0, np.pi / 4, 0.7853982, 119.99999, 120.0, 1e9, -3e38], np.float32)])
    s = np.empty_like(a)
    c = np.empty_like(a)
    from box2d_rs_b200.lib import check
    check(ctx.L, ctx.L.b2gpu_debug_sincos(ctx.h, a.ctypes.data, s.ctypes.data, c.ctypes.data, a.size))
    rs, rc = b2o.sincosf(a)
    assert np.array_equal(s.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), rc.view(np.uint32))


@pytest.mark.parametrize("name", list(SCENES))
def test_free_running_single_world(name, ctx):
    """B2world mirror on the GPU vs oracle, stepping freely from the same scene: bit-identical state."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 3 or i % 25 == 24 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()


@pytest.mark.parametrize("name", ["pyramid", "mixed300", "addpair2000"])
def test_teacher_forced_steps(name, ctx):
    """upload oracle state S_n, step both once, compare — for many n."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    batch = wg.batch(2, lane_block=1)
    for i in range(min(steps, 120)):
        if i % 6 == 0:
            batch.upload_world(1, wo.snapshot())
            batch.step(scenes.DT, 8, 3)
            wo.step(scenes.DT, 8, 3)
            bad = parity.compare_snapshots(wo.snapshot(), batch.download_world(1)) + \
                parity.compare_stats(wo.get_stats(), batch.stats()[1])
            assert bad == [], "step %d: %s" % (i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    batch.close()
    wg.close()


@pytest.mark.parametrize("n_worlds,lane_block", [(70, 32), (5, 4), (33, 0)])
def test_batch_of_different_worlds(n_worlds, lane_block, ctx):
    """Worlds of one batch evolve independently: perturbed Pyramid worlds vs per-world oracle clones."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("pyramid", ctx)
    batch = wg.batch(n_worlds, lane_block=lane_block)
    picks = sorted(set([0, 1, n_worlds // 2, n_worlds - 1]))
    oracles = {}
    for w in picks:
        o = wo.clone()
        o.body(211).set_transform((3.6875 + 0.01 * (w % 7) - 0.03, 24.5 + 0.5 * (w % 3)), 0.05 * (w % 5))
        batch.upload_world(w, o.snapshot())
        oracles[w] = o
    for i in range(60):
        batch.step(scenes.DT, 8, 3)
        for o in oracles.values():
            o.step(scenes.DT, 8, 3)
    for w, o in oracles.items():
        bad = parity.compare_snapshots(o.snapshot(), batch.download_world(w))
        assert bad == [], "world %d: %s" % (w, bad[:6])
    state = batch.body_state()
    for w, o in oracles.items():
        assert np.array_equal(state[w].view(np.uint32), o.body_state().view(np.uint32))
    batch.close()
    wg.close()


@pytest.mark.parametrize("name,steps", [("hello_world", 90), ("mixed300", 260), ("addpair2000", 120), ("pile400", 120), ("variety", 400), ("sensors", 300)])
def test_batch_replicas_shared_memory_solver(name, steps, ctx):
    """32-world memory blocks run the shared-memory Gauss-Seidel kernels (resident ring for tiny islands,
    streaming ring otherwise, many islands per world): replicas must match the oracle bit for bit."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair(name, ctx)
    batch = wg.batch(40)
    for i in range(steps):
        batch.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
        if i % 20 == 19 or i == steps - 1:
            for w in (0, 39):
                bad = parity.compare_snapshots(wo.snapshot(), batch.download_world(w)) + \
                    parity.compare_stats(wo.get_stats(), batch.stats()[w])
                assert bad == [], "step %d world %d: %s" % (i, w, bad[:6])
    batch.close()
    wg.close()


@pytest.mark.parametrize("solver", ["generic", "one_stream", "no_graph"])
@pytest.mark.parametrize("name,steps", [("mixed300", 200), ("variety", 300), ("sensors", 200)])
def test_alternative_solver_kernels(name, steps, solver, ctx):
    """The one-lane-per-world shared-memory kernels and the generic global-memory stages stay available
    (worlds too large for the level-scheduled kernels): same bits as the oracle."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair(name, ctx)
    batch = wg.batch(34, solver=solver)
    for i in range(steps):
        batch.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
    bad = parity.compare_snapshots(wo.snapshot(), batch.download_world(33)) + parity.compare_stats(wo.get_stats(), batch.stats()[33])
    assert bad == [], bad[:6]
    batch.close()
    wg.close()


def test_full_size_batch_properties(ctx):
    """BASELINE config 3 at full size (4096 Pyramid worlds): size-independent properties — replicas stay
    bit-identical to each other and to the oracle; a world pushed by an external force diverges alone."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("pyramid", ctx)
    n = 4096
    batch = wg.batch(n, max_contacts=1024)
    forces = np.zeros((n, batch.body_count, 3), np.float32)
    forces[1234, 211, 0] = 400.0
    for i in range(30):
        batch.set_forces(forces)
        batch.step(scenes.DT, 8, 3)
        wo.step(scenes.DT, 8, 3)
    state = batch.body_state()
    ref = wo.body_state()
    same = (state.view(np.uint32) == ref.view(np.uint32)[None]).all(axis=(1, 2))
    assert same.sum() == n - 1 and not same[1234]
    st = batch.stats()
    assert (st["status"] == 0).all()
    assert (st["contacts"] == int(wo.get_stats()["contacts"]))[np.arange(n) != 1234].all()
    batch.close()
    wg.close()


@pytest.mark.parametrize("name", ["pyramid", "mixed300", "pile400", "addpair2000", "hello_world", "variety", "sensors", "terrain"])
def test_gpu_matches_golden(name, ctx):
    """The committed fixtures (tests/golden, produced by the oracle) reproduced by the CUDA path, in a
    40-world batch (shared-memory solver and island kernels), bit for bit."""
    import os
    import sys
    from conftest import ROOT, SCENES as SC
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    from box2d_rs_b200 import scenes, world
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    recipe, gravity, _ = SC[name]
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    batch = wg.batch(40)

    class View:
        def body_state(self):
            return batch.body_state(39, 1)[0]

        def snapshot(self):
            return batch.download_world(39)

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


@pytest.mark.parametrize("allow_sleep", [True, False])
def test_config1_pyramid_1000_steps(allow_sleep, ctx):
    """BASELINE configs[0]: testbed Pyramid, 1000 steps, dt = 1/60, 8/3 iterations, continuous off —
    free-running in a 64-world batch, compared with the oracle every 100 steps, bit for bit."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("pyramid", ctx)
    wo.set_allow_sleeping(allow_sleep)
    wg.set_allow_sleeping(allow_sleep)
    batch = wg.batch(64, max_contacts=1024)
    for i in range(10):
        batch.step(scenes.DT, 8, 3, 100)
        for _ in range(100):
            wo.step(scenes.DT, 8, 3)
        bad = parity.compare_snapshots(wo.snapshot(), batch.download_world(63)) + \
            parity.compare_stats(wo.get_stats(), batch.stats()[63])
        assert bad == [], "step %d: %s" % (100 * (i + 1), bad[:6])
    state = batch.body_state()
    assert (state.view(np.uint32) == state[0].view(np.uint32)[None]).all()
    batch.close()
    wg.close()


@pytest.mark.parametrize("first,count,batch_mode", [(0, 40, True), (1000, 25, False)])
def test_differential_fuzz(first, count, batch_mode, ctx):
    """tools/fuzz_parity.py: random scenes, flags, iteration counts, dt = 0 steps and mid-run edits; the CUDA
    path (batched: shared-memory island / solver kernels; single world: generic stages) vs the oracle.
    Seeds 12 and 39 once caught stale lazily-materialised contact ISLAND bits after a collide-only step."""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_parity
    from box2d_rs_b200 import world
    fails = []
    for seed in range(first, first + count):
        r = fuzz_parity.run_seed(seed, lambda g: world.B2world(g, ctx=ctx), 128, batch_mode)
        if r not in (None, "skip"):
            fails.append((seed, r))
    assert fails == []


def test_step_host_overlapped_copies_match_plain_calls(ctx):
    """b2gpu_batch_step_host overlaps the force upload and the state download with stages that do not touch
    them; its results must equal set_forces + step + get_body_state, and the oracle."""
    from box2d_rs_b200 import scenes
    wo, wg, _ = _pair("pyramid", ctx)
    n = 64
    a = wg.batch(n, max_contacts=1024)
    b = wg.batch(n, max_contacts=1024)
    rng = np.random.default_rng(5)
    state = np.zeros((n, a.body_count, 8), np.float32)
    for i in range(40):
        forces = np.zeros((n, a.body_count, 3), np.float32)
        forces[:, 2:, 0] = rng.uniform(-30.0, 30.0, (n, a.body_count - 2)).astype(np.float32)
        a.step_host(forces, state, scenes.DT, 8, 3, 1)
        b.set_forces(forces)
        b.step(scenes.DT, 8, 3)
        assert np.array_equal(state.view(np.uint32), b.body_state().view(np.uint32)), "step %d" % i
        for k in range(2, a.body_count):
            wo.body(k).apply_force_to_center((float(forces[n - 1, k, 0]), 0.0), wake=False)
        wo.step(scenes.DT, 8, 3)
    assert np.array_equal(state[n - 1].view(np.uint32), wo.body_state().view(np.uint32))
    assert parity.compare_snapshots(wo.snapshot(), a.download_world(n - 1)) == []
    a.close()
    b.close()
    wg.close()


# ---------------------------------------------------------------------------------------------------------
# large-world mode (b2g_large.h): data-parallel broadphase / destruction / islands for one big world
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,every", [("pyramid", 3), ("mixed300", 3), ("pile400", 2), ("variety", 2), ("sensors", 3),
                                        ("addpair2000", 5), ("terrain", 3)])
def test_large_mode_teacher_forced(name, every, ctx):
    """Every step of the large-world mode is the oracle's step of the same state: upload S_n, step both once,
    compare everything bit for bit (contacts created in that step as a set: they are appended in LBVH order)."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    bt = wg.batch(1, lane_block=1, solver='large')
    for i in range(steps):
        if i % every == 0:
            bt.upload_world(0, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(0), wo.get_stats(), bt.stats()[0])
            assert bad == [], "step %d: %s" % (i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    bt.close()
    wg.close()


@pytest.mark.parametrize("name,n_steps", [("pile400", 200), ("addpair2000", 150), ("mixed300", 200)])
def test_large_mode_free_running_is_deterministic(name, n_steps, ctx):
    """Free-running, the large-world mode on the GPU (atomics, union-find hooking, refit arrival order, CUB
    scans and sorts) and its sequential host simulation (test infrastructure) produce the same bits."""
    from box2d_rs_b200 import batch as batch_mod, scenes, world
    from conftest import HOSTSIM_SO
    recipe, gravity, _ = SCENES[name]
    hctx = batch_mod.Context(0, lib_path=HOSTSIM_SO)
    ws = []
    for c in (ctx, hctx):
        w = world.B2world(gravity, ctx=c)
        recipe(scenes, w)
        w.set_large_mode(True)
        ws.append(w)
    for i in range(n_steps):
        for w in ws:
            w.step(scenes.DT, 8, 3)
        if i % 50 == 49 or i == n_steps - 1:
            bad = parity.compare_snapshots(ws[1].snapshot(), ws[0].snapshot()) + parity.compare_stats(ws[1].get_stats(), ws[0].get_stats())
            assert bad == [], "step %d: %s" % (i, bad[:6])
    assert int(ws[0].get_stats()["status"]) == 0
    for w in ws:
        w.close()
    hctx.close()


@pytest.mark.parametrize("name", ["pyramid", "mixed300", "pile400", "addpair2000", "variety", "sensors", "terrain"])
def test_large_mode_exact_order_free_running(name, ctx):
    """Large-world mode with the replica tree kept (flag 2): free-running, the whole snapshot — tree included —
    is bit-identical to the oracle."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    wg.set_large_mode(2)
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 2 or i % 50 == 49 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + \
                [b for b in parity.compare_stats(wo.get_stats(), wg.get_stats()) if "island_bodies" not in b]
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()


# ---- level-scheduled sweeps of giant islands (b2g_levels.h): one CTA per island, a barrier per dependency level
@pytest.mark.parametrize("name,every", [("pyramid", 3), ("mixed300", 3), ("pile400", 2), ("variety", 2), ("addpair2000", 3), ("terrain", 3)])
def test_level_scheduled_islands_teacher_forced(name, every, ctx):
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    bt = wg.batch(1, lane_block=1, solver='large')
    bt.set_level_threshold(6)
    levels = 0
    for i in range(steps):
        if i % every == 0:
            bt.upload_world(0, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            st = bt.stats()[0]
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(0), wo.get_stats(), st)
            assert bad == [], "step %d: %s" % (i, bad[:6])
            levels = max(levels, int(st["solver_levels"]))
        else:
            wo.step(scenes.DT, 8, 3)
    assert levels > 0
    bt.close()
    wg.close()


@pytest.mark.parametrize("name", ["pyramid", "pile400", "mixed300", "variety"])
def test_level_scheduled_islands_free_running(name, ctx):
    """Mode 2 with every island of >= 4 contacts swept level by level: bit-identical to the oracle free-running."""
    from box2d_rs_b200 import scenes
    wo, wg, steps = _pair(name, ctx)
    wg.set_large_mode(2)
    wg.set_level_threshold(4)
    levels = 0
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 2 or i % 50 == 49 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + \
                [b for b in parity.compare_stats(wo.get_stats(), wg.get_stats()) if "island_bodies" not in b]
            assert bad == [], "step %d: %s" % (i, bad[:6])
            levels = max(levels, int(wg.get_stats()["solver_levels"]))
    assert levels > 0
    wg.close()


@pytest.mark.parametrize("warm", [True, False])
def test_level_scheduled_shallow_islands(warm, ctx):
    """Islands of three levels (a stack of three boxes), with and without warm starting: 9 or 8 passes over 3 levels.  (With 8
    passes the consumers' last published position used to sit in the previous pass, and the producer warp waited for room
    that never came.)"""
    from box2d_rs_b200 import abi, scenes, world
    from oracle import b2o

    def build(w):
        ground = w.create_body(abi.BodyDef())
        ground.create_fixture_by_shape(w.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
        box = w.shapes.polygon_box(0.5, 0.5)
        for s in range(6):
            for i in range(3 + s % 3):
                b = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY, position=(-20.0 + 6.0 * s, 0.51 + 1.02 * i), allow_sleep=0))
                b.create_fixture_by_shape(box, 1.0)

    wo = b2o.B2world((0.0, -10.0))
    build(wo)
    wg = world.B2world((0.0, -10.0), ctx=ctx)
    build(wg)
    for w in (wo, wg):
        w.set_warm_starting(warm)
    wg.set_large_mode(2)
    wg.set_level_threshold(3)
    levels = 0
    for i in range(60):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        levels = max(levels, int(wg.get_stats()["solver_levels"]))
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    assert levels >= 3
    wg.close()


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs[1], [3], [4] at FULL size (10k mixed, 100k pile, AddPair-20k): teacher-forced single steps
# ---------------------------------------------------------------------------------------------------------
FULL_SIZE = {
    # name: (recipe, gravity, oracle steps at which one teacher-forced GPU step is compared)
    "mixed10k": (lambda s, w: s.mixed(w, n=10000), (0.0, -10.0), (0, 25, 60, 125)),  # step 125: an island of > 1024 contacts (level-scheduled sweep)
    "pile100k": (lambda s, w: s.pile(w, n=100000), (0.0, -10.0), (0, 12, 30)),
    "addpair20k": (lambda s, w: s.add_pair(w, n=20000), (0.0, 0.0), (0, 26, 40)),
}


@pytest.mark.parametrize("mode", ["large", "large_exact"])
@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_size_teacher_forced(name, mode, ctx):
    """The oracle runs the full-size scene; at the listed steps its whole state S_n is uploaded, both engines step
    once, and everything is compared bit for bit: in mode 1 the contacts created in that step as a set (LBVH
    append order), in mode 2 (reference contact order) the complete snapshot including the replica tree."""
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    recipe, gravity, at = FULL_SIZE[name]
    wo = b2o.B2world(gravity)
    recipe(scenes, wo)
    wg = world.B2world(gravity, ctx=ctx)
    recipe(scenes, wg)
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    bt = wg.batch(1, lane_block=1, solver=mode)
    for i in range(max(at) + 1):
        if i in at:
            bt.upload_world(0, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            got, so, sg = bt.download_world(0), wo.get_stats(), bt.stats()[0]
            if mode == "large":
                bad = parity.compare_large_step(wo.snapshot(), got, so, sg)
            else:
                bad = parity.compare_snapshots(wo.snapshot(), got) + \
                    [b for b in parity.compare_stats(so, sg) if "island_bodies" not in b]
            assert bad == [], "%s step %d: %s" % (name, i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    assert int(wo.get_stats()["contacts"]) > 1000
    bt.close()
    wg.close()
