"""CPU tests (no GPU): the oracle against every known-answer test the reference holds for this path
(SURVEY.md §4 / §8c): tests/test.rs hello_world, tests/world_test.rs begin_contact,
tests/collision_test.rs polygon_mass_data, tests/math_test.rs sweep — plus the libm pin of the
restated sinf/cosf and the committed golden fixtures.  The same known answers are asserted for the
product's host-side world builder (box2d_rs_b200/csrc/b2g_world.cu) where no device is needed."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from conftest import HOSTSIM_SO, ROOT, SCENES

EPS = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module")
def b2o(built):
    from oracle import b2o as m
    return m


def test_hello_world_known_answer(b2o):
    """tests/test.rs:25-102: 1x1 box (density 1, friction 0.3) dropped from y = 4 on a 50x10 static box."""
    from box2d_rs_b200 import scenes
    w = b2o.B2world((0.0, -10.0))
    body = scenes.hello_world(w)
    for _ in range(60):
        w.step(scenes.DT, 6, 2)
    x, y = body.get_position()
    assert abs(x) < 0.01 and abs(y - 1.01) < 0.01 and abs(body.get_angle()) < 0.01


def test_begin_contact(b2o):
    """tests/world_test.rs:50-88: no contact at distance 100; after set_transform to distance 1 one step
    creates the contact (m_new_contacts -> find_new_contacts) and it is touching (begin_contact fires)."""
    from box2d_rs_b200 import abi, scenes
    w = b2o.B2world((0.0, -10.0))
    circle = w.shapes.circle(5.0)
    a = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY))
    b = w.create_body(abi.BodyDef(type=abi.DYNAMIC_BODY))
    a.create_fixture_by_shape(circle, 0.0)
    b.create_fixture_by_shape(circle, 0.0)
    a.set_transform((0.0, 0.0), 0.0)
    b.set_transform((100.0, 0.0), 0.0)
    w.step(scenes.DT, 6, 2)
    assert w.get_contact_count() == 0
    b.set_transform((1.0, 0.0), 0.0)
    w.step(scenes.DT, 6, 2)
    assert w.get_contact_count() == 1
    snap = w.snapshot()
    assert int(snap.contacts["flags"][0]) & 0x2  # TOUCHING: begin_contact would have fired


def _mass_checks(shapes):
    center = (100.0, -50.0)
    hx, hy, angle1 = 0.5, 1.5, 0.25
    abs_tol = rel_tol = 2.0 * EPS
    p1 = shapes.polygon_box(hx, hy, center, angle1)
    assert abs(p1.centroid[0] - center[0]) < abs_tol + rel_tol * abs(center[0])
    assert abs(p1.centroid[1] - center[1]) < abs_tol + rel_tol * abs(center[1])
    p2 = shapes.polygon([(center[0] - hx, center[1] - hy), (center[0] + hx, center[1] - hy),
                         (center[0] - hx, center[1] + hy), (center[0] + hx, center[1] + hy)])
    assert abs(p2.centroid[0] - center[0]) < abs_tol + rel_tol * abs(center[0])
    assert abs(p2.centroid[1] - center[1]) < abs_tol + rel_tol * abs(center[1])
    mass = 4.0 * hx * hy
    inertia = (mass / 3.0) * (hx * hx + hy * hy) + mass * (center[0] ** 2 + center[1] ** 2)
    for p in (p1, p2):
        md = shapes.compute_mass(p, 1.0)
        assert abs(md.center_x - center[0]) < abs_tol + rel_tol * abs(center[0])
        assert abs(md.center_y - center[1]) < abs_tol + rel_tol * abs(center[1])
        assert abs(md.mass - mass) < 20.0 * (abs_tol + rel_tol * mass)
        assert abs(md.inertia - inertia) < 40.0 * (abs_tol + rel_tol * inertia)


def test_polygon_mass_data_oracle(b2o):
    """tests/collision_test.rs:29-81 on the oracle's restated setup geometry."""
    _mass_checks(b2o.Shapes)


def test_polygon_mass_data_product(built):
    """The same known answer on the product's host-side builder (no device needed for shapes)."""
    from box2d_rs_b200 import lib, world
    _mass_checks(world.Shapes(lib.load()))


def test_sweep_matches_libm(b2o):
    """tests/math_test.rs:25-49: B2Sweep::get_transform endpoints equal f32::sin/cos exactly."""
    sweep = (C.c_float * 8)(0.0, 0.0, -2.0, 4.0, 3.0, 8.0, 0.5, 5.0)
    xf = (C.c_float * 4)()
    for beta, c, a in ((0.0, (-2.0, 4.0), 0.5), (1.0, (3.0, 8.0), 5.0)):
        b2o.lib().b2o_sweep_get_transform(sweep, beta, xf)
        s_ref, c_ref = b2o.sincosf(np.array([a], np.float32))
        assert (xf[0], xf[1]) == c
        assert np.float32(xf[2]) == s_ref[0] and np.float32(xf[3]) == c_ref[0]


def test_restated_sincos_matches_libm_on_host(b2o):
    """The restated glibc sinf/cosf (b2g_math.h, the code the device runs) against the host libm:
    2^22 random bit patterns + dense ranges.  (Exhaustive 2^32 check: oracle/sincosf_check.c.)"""
    from box2d_rs_b200 import batch
    from box2d_rs_b200.lib import check
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    rng = np.random.default_rng(11)
    a = rng.integers(0, 2**32, size=1 << 22, dtype=np.uint64).astype(np.uint32).view(np.float32)
    a = a[np.isfinite(a)]
    a = np.concatenate([a, rng.uniform(-130.0, 130.0, 1 << 20).astype(np.float32),
                        np.array([0.0, -0.0, 0.7853981, 0.7853982, 119.99999, 120.0, 1e9, -3e38, 1e-40], np.float32)])
    s = np.empty_like(a)
    c = np.empty_like(a)
    check(ctx.L, ctx.L.b2gpu_debug_sincos(ctx.h, a.ctypes.data, s.ctypes.data, c.ctypes.data, a.size))
    rs, rc = b2o.sincosf(a)
    assert np.array_equal(s.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), rc.view(np.uint32))
    ctx.close()


@pytest.mark.parametrize("name", ["pyramid", "hello_world", "mixed300", "pile400", "addpair2000", "variety", "sensors", "terrain"])
def test_oracle_matches_golden(name, b2o):
    """The committed fixtures (tests/golden/make_golden.py) pin the oracle's trajectories bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    from box2d_rs_b200 import scenes
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    recipe, gravity, _ = SCENES[name]
    w = b2o.B2world(gravity)
    recipe(scenes, w)
    got = make_golden.record(w, [int(s) for s in g["steps"]], lambda: w.step(scenes.DT, scenes.VEL_ITERS, scenes.POS_ITERS))
    for k, v in got.items():
        ref = g[k]
        if k.startswith("state"):
            assert np.array_equal(ref.view(np.uint32), v.view(np.uint32)), k
        else:
            assert np.array_equal(ref, v), k


def test_pyramid_structure(b2o):
    """testbed Pyramid = 212 bodies, 211 proxies; 590 contacts (400 touching, one island of
    211 bodies) once settled; it falls asleep with sleeping allowed."""
    from box2d_rs_b200 import scenes
    w = b2o.B2world((0.0, -10.0))
    scenes.pyramid(w)
    snap = w.snapshot()
    assert snap.n.body_count == 212 and snap.n.proxy_count == 211
    for _ in range(120):
        w.step(scenes.DT, 8, 3)
    st = w.get_stats()
    assert (int(st["contacts"]), int(st["touching"]), int(st["islands"]), int(st["island_bodies"])) == (590, 400, 1, 211)
    for _ in range(200):
        w.step(scenes.DT, 8, 3)
    assert int(w.get_stats()["awake_bodies"]) == 0
