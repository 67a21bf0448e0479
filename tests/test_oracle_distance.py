"""Known answers for the oracle's restatement of the GJK distance query (oracle/b2o_distance.hpp follows
src/private/collision/b2_distance.rs; the reference ships no test for it, so the answers are analytic):
distances between circles, boxes and edges with and without radii, witness points, overlap, symmetry, and the
b2_test_overlap threshold (distance < 10 * epsilon) that decides `touching` of sensor contacts."""
import math

import numpy as np
import pytest

EPS = float(np.finfo(np.float32).eps)
POLY_R = 0.01  # b2_polygon_radius = 2 * linear_slop


@pytest.fixture(scope="module")
def o():
    from oracle import b2o
    return b2o


def test_circle_circle(o):
    a, b = o.Shapes.circle(0.5), o.Shapes.circle(0.25)
    pa, pb, d, it = o.shape_distance(a, (0, 0, 0), b, (2, 0, 0))
    assert d == pytest.approx(1.25, rel=1e-6)
    assert pa == pytest.approx((0.5, 0.0), abs=1e-6) and pb == pytest.approx((1.75, 0.0), abs=1e-6)
    _, _, d0, _ = o.shape_distance(a, (0, 0, 0), b, (2, 0, 0), use_radii=False)
    assert d0 == pytest.approx(2.0, rel=1e-6)
    assert it <= 2


def test_box_box_face_and_corner(o):
    a, b = o.Shapes.polygon_box(0.5, 0.5), o.Shapes.polygon_box(0.5, 0.5)
    _, _, d, _ = o.shape_distance(a, (0, 0, 0), b, (3, 0, 0))
    assert d == pytest.approx(2.0 - 2 * POLY_R, rel=1e-6)
    _, _, d, _ = o.shape_distance(a, (0, 0, 0), b, (3, 0, 0), use_radii=False)
    assert d == pytest.approx(2.0, rel=1e-6)
    # corner to corner along the diagonal
    pa, pb, d, _ = o.shape_distance(a, (0, 0, 0), b, (3, 3, 0), use_radii=False)
    assert d == pytest.approx(math.sqrt(8.0), rel=1e-6)
    assert pa == pytest.approx((0.5, 0.5), abs=1e-6) and pb == pytest.approx((2.5, 2.5), abs=1e-6)
    # a box rotated by 45 degrees points a corner at the other box's face
    _, _, d, _ = o.shape_distance(a, (0, 0, 0), b, (3, 0, math.pi / 4), use_radii=False)
    assert d == pytest.approx(3.0 - 0.5 - math.sqrt(0.5), rel=1e-5)


def test_box_circle_and_edge_circle(o):
    box, c = o.Shapes.polygon_box(1.0, 1.0), o.Shapes.circle(0.5)
    _, _, d, _ = o.shape_distance(box, (0, 0, 0), c, (3, 3, 0))
    assert d == pytest.approx(math.sqrt(8.0) - 0.5 - POLY_R, rel=1e-6)
    e = o.Shapes.edge_two_sided((-1.0, 0.0), (1.0, 0.0))
    c2 = o.Shapes.circle(0.3)
    pa, pb, d, _ = o.shape_distance(e, (0, 0, 0), c2, (0.25, 1.0, 0))
    assert d == pytest.approx(1.0 - 0.3 - POLY_R, rel=1e-6)
    assert pa == pytest.approx((0.25, POLY_R), abs=1e-6) and pb == pytest.approx((0.25, 0.7), abs=1e-6)
    # beyond the end of the segment the closest feature is the end point
    _, _, d, _ = o.shape_distance(e, (0, 0, 0), c2, (2.0, 1.0, 0), use_radii=False)
    assert d == pytest.approx(math.sqrt(2.0), rel=1e-6)


def test_overlap_and_threshold(o):
    a, b = o.Shapes.polygon_box(0.5, 0.5), o.Shapes.polygon_box(0.5, 0.5)
    _, _, d, _ = o.shape_distance(a, (0, 0, 0), b, (0.6, 0.2, 0.3))
    assert d == 0.0
    assert o.test_overlap_shapes(a, (0, 0, 0), b, (0.6, 0.2, 0.3))
    assert not o.test_overlap_shapes(a, (0, 0, 0), b, (1.5, 0, 0))
    # touching through the polygon skins counts as overlap, a hair more does not
    assert o.test_overlap_shapes(a, (0, 0, 0), b, (1.0 + 2 * POLY_R, 0, 0))
    assert not o.test_overlap_shapes(a, (0, 0, 0), b, (1.0 + 2 * POLY_R + 1e-4, 0, 0))
    c = o.Shapes.circle(0.5)
    assert o.test_overlap_shapes(c, (0, 0, 0), c, (1.0, 0, 0))
    assert not o.test_overlap_shapes(c, (0, 0, 0), c, (1.0 + 100 * EPS, 0, 0))


def test_symmetry_and_polygon_inside(o):
    rng = np.random.default_rng(7)
    tri = o.Shapes.polygon([(0.0, 0.0), (1.0, 0.0), (0.2, 0.8)])
    hexa = o.Shapes.polygon([(0.5 * math.cos(i * math.pi / 3), 0.4 * math.sin(i * math.pi / 3)) for i in range(6)])
    for _ in range(200):
        xa = (rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-3, 3))
        xb = (rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-3, 3))
        pa, pb, d, it = o.shape_distance(tri, xa, hexa, xb)
        qb, qa, d2, _ = o.shape_distance(hexa, xb, tri, xa)
        assert it <= 20
        assert d == pytest.approx(d2, abs=2e-6)
        if d > 0:  # the witness points realise the distance
            assert math.hypot(pa[0] - pb[0], pa[1] - pb[1]) == pytest.approx(d, abs=2e-6)
    small = o.Shapes.circle(0.05)
    assert o.test_overlap_shapes(hexa, (0, 0, 0), small, (0.1, 0.05, 0))  # circle strictly inside the polygon


def test_chain_child(o):
    chain = o.Shapes.chain([(0.0, 0.0), (1.0, 0.0), (2.0, 1.0)], (-1.0, 0.0), (3.0, 1.0))
    c = o.Shapes.circle(0.1)
    _, _, d0, _ = o.shape_distance(chain, (0, 0, 0), c, (0.5, 0.5, 0), use_radii=False, index_a=0)
    assert d0 == pytest.approx(0.5, rel=1e-6)
    _, _, d1, _ = o.shape_distance(chain, (0, 0, 0), c, (0.5, 0.5, 0), use_radii=False, index_a=1)
    assert d1 == pytest.approx(math.sqrt(0.5), rel=1e-6)  # closest feature of child 1 is its first vertex (1, 0)
