"""Scene recipes of BASELINE.json's configs, written against the mirrored
reference API (world.create_body / body.create_fixture / world.shapes.*), so the same recipe
drives the GPU engine and — in tests — the CPU oracle.

Randomised scenes use SplitMix64 (the reference's testbed uses unseeded rand::thread_rng,
examples/testbed/test.rs:27-36); seed = 0xB2D + config id.
"""
import math

import numpy as np

from . import abi
from .abi import BodyDef, FixtureDef

DT = float(np.float32(1.0) / np.float32(60.0))
VEL_ITERS, POS_ITERS = 8, 3


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def uniform(self, lo, hi):
        return lo + (hi - lo) * ((self.next() >> 40) / float(1 << 24))


def f32(x):
    return float(np.float32(x))


def hello_world(world):
    """tests/test.rs:25-102 — returns the dynamic body."""
    ground = world.create_body(BodyDef(position=(0.0, -10.0)))
    ground.create_fixture_by_shape(world.shapes.polygon_box(50.0, 10.0), 0.0)
    body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 4.0)))
    body.create_fixture(FixtureDef(density=1.0, friction=0.3), world.shapes.polygon_box(1.0, 1.0))
    return body


def pyramid(world, count=20, testbed_ground_body=True):
    """examples/testbed/tests/pyramid.rs:52-88 preceded by the testbed's empty static body
    (examples/testbed/test.rs:181-182).  212 bodies, 211 proxies for count=20."""
    if testbed_ground_body:
        world.create_body(BodyDef())
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
    box = world.shapes.polygon_box(0.5, 0.5)
    x = np.array([-7.0, 0.75], np.float32)
    delta_x = np.array([0.5625, 1.25], np.float32)
    delta_y = np.array([1.125, 0.0], np.float32)
    bodies = []
    for i in range(count):
        y = x.copy()
        for _ in range(i, count):
            b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(float(y[0]), float(y[1]))))
            b.create_fixture_by_shape(box, 5.0)
            bodies.append(b)
            y = (y + delta_y).astype(np.float32)
        x = (x + delta_x).astype(np.float32)
    return bodies


def _container(world, half_width, height):
    """Static open box made of three two-sided edges (floor + walls)."""
    ground = world.create_body(BodyDef())
    hw = float(half_width)
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-hw, 0.0), (hw, 0.0)), 0.0)
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-hw, 0.0), (-hw, float(height))), 0.0)
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((hw, 0.0), (hw, float(height))), 0.0)
    return ground


def _polygon_archetypes(world):
    """The five archetypes of examples/testbed/tests/polygon_shapes.rs:106-188."""
    tri = world.shapes.polygon([(-0.5, 0.0), (0.5, 0.0), (0.0, 1.5)])
    thin = world.shapes.polygon([(-0.1, 0.0), (0.1, 0.0), (0.0, 1.5)])
    w = 1.0
    b = w / (2.0 + math.sqrt(2.0))
    s = math.sqrt(2.0) * b
    octagon = world.shapes.polygon([(0.5 * s, 0.0), (0.5 * w, b), (0.5 * w, b + s), (0.5 * s, w), (-0.5 * s, w),
                                    (-0.5 * w, b + s), (-0.5 * w, b), (-0.5 * s, 0.0)])
    box = world.shapes.polygon_box(0.5, 0.5)
    circle = world.shapes.circle(0.5)
    return [tri, thin, octagon, box, circle]


def mixed(world, n=10000, seed=0xB2D + 2, width=100.0):
    """Config 2: n mixed polygons + circles on a jittered grid above a static container."""
    rng = SplitMix64(seed)
    _container(world, width / 2.0, 60.0 if n >= 1000 else 20.0)
    shapes = _polygon_archetypes(world)
    cols = max(1, int(width / 2.0) - 2)
    bodies = []
    for k in range(n):
        col, row = k % cols, k // cols
        px = f32(-width / 2.0 + 2.5 + 2.0 * col + rng.uniform(-0.2, 0.2))
        py = f32(2.0 + 2.0 * row + rng.uniform(-0.2, 0.2))
        ang = f32(rng.uniform(-math.pi, math.pi))
        kind = int(rng.next() % 5)
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py), angle=ang))
        b.create_fixture(FixtureDef(density=1.0, friction=0.3), shapes[kind])
        bodies.append(b)
    return bodies


def pile(world, n=100000, seed=0xB2D + 4, width=400.0):
    """Config 4: n/2 circles r=0.125 (friction 0.1) + n/2 boxes half-extent 0.125 (friction 0.3)."""
    rng = SplitMix64(seed)
    _container(world, width / 2.0, 100.0)
    circle = world.shapes.circle(0.125)
    box = world.shapes.polygon_box(0.125, 0.125)
    pitch = 0.3
    cols = max(1, int((width - 2.0) / pitch))
    bodies = []
    for k in range(n):
        col, row = k % cols, k // cols
        px = f32(-width / 2.0 + 1.0 + pitch * col + rng.uniform(-0.02, 0.02))
        py = f32(0.3 + pitch * row + rng.uniform(-0.02, 0.02))
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py)))
        if k % 2 == 0:
            b.create_fixture(FixtureDef(density=1.0, friction=0.1), circle)
        else:
            b.create_fixture(FixtureDef(density=1.0, friction=0.3), box)
        bodies.append(b)
    return bodies


def add_pair(world, n=20000, seed=0xB2D + 5):
    """Config 5: examples/testbed/tests/add_pair.rs:50-89 scaled. The world must have zero gravity."""
    rng = SplitMix64(seed)
    circle = world.shapes.circle(0.1)
    bodies = []
    for _ in range(n):
        px, py = f32(rng.uniform(-60.0, 0.0)), f32(rng.uniform(-10.0, 20.0))
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py)))
        b.create_fixture_by_shape(circle, 0.01)
        bodies.append(b)
    box = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-100.0, 5.0), bullet=1))
    box.create_fixture_by_shape(world.shapes.polygon_box(1.5, 1.5), 1.0)
    box.set_linear_velocity((100.0, 0.0))
    bodies.append(box)
    return bodies


def variety(world, seed=0xB2D + 9):
    """Edge-case scene for the parity tests: a chain-shape bowl (one-sided edges with ghost vertices), a
    kinematic paddle, restitution, collision filtering by group and by category/mask, polygons of 3..8
    vertices, circles, a multi-fixture body, a fixed-rotation body, damping and gravity scale."""
    rng = SplitMix64(seed)
    ground = world.create_body(BodyDef())
    bowl = [(-12.0, 8.0), (-10.0, 1.0), (-5.0, 0.0), (0.0, -0.5), (5.0, 0.0), (10.0, 1.0), (12.0, 8.0)]
    bowl = [(x, y) for x, y in reversed(bowl)]  # counter-clockwise so the solid side faces the inside
    ground.create_fixture(FixtureDef(friction=0.6), world.shapes.chain(bowl, (13.0, 9.0), (-13.0, 9.0)))
    paddle = world.create_body(BodyDef(type=abi.KINEMATIC_BODY, position=(-6.0, 3.0), angle=0.3,
                                       linear_velocity=(1.5, 0.0), angular_velocity=0.7))
    paddle.create_fixture_by_shape(world.shapes.polygon_box(2.0, 0.2), 0.0)
    bodies = []
    for k in range(60):
        px = f32(-8.0 + 16.0 * ((k * 7) % 60) / 60.0 + rng.uniform(-0.1, 0.1))
        py = f32(4.0 + 0.9 * (k // 6) + rng.uniform(-0.1, 0.1))
        kind = k % 6
        bd = BodyDef(type=abi.DYNAMIC_BODY, position=(px, py), angle=f32(rng.uniform(-3.0, 3.0)))
        if kind == 4:
            bd.fixed_rotation = 1
        if kind == 5:
            bd.linear_damping, bd.angular_damping, bd.gravity_scale = 0.3, 0.2, 0.5
        b = world.create_body(bd)
        fd = FixtureDef(density=1.0 + 0.5 * kind, friction=0.1 * (1 + kind), restitution=0.0)
        if kind == 0:
            fd.restitution = 0.6
            shape = world.shapes.circle(0.3 + 0.02 * (k % 5))
        elif kind == 1:
            nv = 3 + (k // 6) % 6
            shape = world.shapes.polygon([(0.45 * math.cos(2 * math.pi * i / nv), 0.35 * math.sin(2 * math.pi * i / nv))
                                          for i in range(nv)])
        elif kind == 2:
            fd.group_index = -3  # never collide with each other
            shape = world.shapes.polygon_box(0.3, 0.5)
        elif kind == 3:
            fd.category_bits, fd.mask_bits = 0x0004, 0xFFFB  # ignore their own category
            shape = world.shapes.circle(0.35)
        elif kind == 4:
            shape = world.shapes.polygon_box(0.4, 0.4, (0.1, -0.05), 0.4)
        else:
            shape = world.shapes.polygon_box(0.5, 0.2)
        b.create_fixture(fd, shape)
        if kind == 5:  # second fixture: multi-fixture mass data, two proxies per body
            b.create_fixture(FixtureDef(density=2.0, friction=0.4), world.shapes.circle(0.2, (0.45, 0.0)))
        bodies.append(b)
    return bodies


def sensors(world, seed=0xB2D + 11):
    """Sensor fixtures of every shape kind (b2_contact.rs(private):149-163: touching = GJK overlap, no
    manifold, no response): a circular and a polygonal sensor zone on the ground body, a one-sided chain
    sensor, a sensor edge, and dynamic bodies that carry a sensor halo next to their solid fixture, raining
    through the zones onto a floor."""
    rng = SplitMix64(seed)
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-12.0, 0.0), (12.0, 0.0)), 0.0)
    ground.create_fixture(FixtureDef(is_sensor=1), world.shapes.circle(2.0, (-5.0, 4.0)))
    ground.create_fixture(FixtureDef(is_sensor=1), world.shapes.polygon_box(2.5, 1.0, (4.0, 3.0), 0.35))
    ground.create_fixture(FixtureDef(is_sensor=1), world.shapes.edge_two_sided((-3.0, 6.0), (3.0, 7.0)))
    ground.create_fixture(FixtureDef(is_sensor=1),
                          world.shapes.chain([(9.0, 2.0), (7.0, 2.5), (5.5, 4.5), (6.0, 7.0)], (10.0, 2.0), (6.5, 8.0)))
    mover = world.create_body(BodyDef(type=abi.KINEMATIC_BODY, position=(-8.0, 2.0), linear_velocity=(2.0, 0.3),
                                      angular_velocity=-0.9))
    mover.create_fixture(FixtureDef(is_sensor=1), world.shapes.polygon([(-1.2, -0.4), (1.0, -0.6), (1.4, 0.5), (-0.2, 0.9)]))
    bodies = []
    for k in range(48):
        px = f32(-9.0 + 18.0 * ((k * 11) % 48) / 48.0 + rng.uniform(-0.15, 0.15))
        py = f32(5.0 + 0.8 * (k // 8) + rng.uniform(-0.1, 0.1))
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py), angle=f32(rng.uniform(-3.0, 3.0))))
        kind = k % 4
        if kind == 0:
            b.create_fixture(FixtureDef(density=1.0, friction=0.3), world.shapes.circle(0.3))
        elif kind == 1:
            b.create_fixture(FixtureDef(density=1.0, friction=0.3), world.shapes.polygon_box(0.35, 0.25))
        elif kind == 2:  # solid core + sensor halo on the same body
            b.create_fixture(FixtureDef(density=2.0, friction=0.5), world.shapes.polygon_box(0.25, 0.25))
            b.create_fixture(FixtureDef(density=0.0, is_sensor=1), world.shapes.circle(0.7))
        else:            # a body made only of a sensor polygon next to a small solid circle
            b.create_fixture(FixtureDef(density=1.5, friction=0.2), world.shapes.circle(0.2, (0.0, -0.2)))
            nv = 3 + (k // 4) % 5
            b.create_fixture(FixtureDef(density=0.0, is_sensor=1),
                             world.shapes.polygon([(0.6 * math.cos(2 * math.pi * i / nv), 0.45 * math.sin(2 * math.pi * i / nv) + 0.3)
                                                   for i in range(nv)]))
        bodies.append(b)
    return bodies


def terrain(world, n=160, segments=1200, seed=0xB2D + 13):
    """Chain shapes at scale: ONE chain fixture of `segments` one-sided edges with ghost
    vertices (a rolling height field, 0.5 m per edge: `segments` proxies, child-edge materialisation per
    contact, b2_chain_shape.rs(private):59-79) plus a closed chain loop as an obstacle; circles and polygons
    of every archetype rain on it, roll downhill and pile up in the valleys."""
    rng = SplitMix64(seed)
    ground = world.create_body(BodyDef())
    half = 0.25 * segments
    pts = []
    for i in range(segments + 1):
        x = half - 0.5 * i  # right to left: the solid side of a one-sided chain is on the right of its direction
        y = 3.0 * math.sin(0.045 * x) + 1.2 * math.sin(0.31 * x + 1.0) + 0.02 * abs(x)
        pts.append((f32(x), f32(y)))
    ground.create_fixture(FixtureDef(friction=0.6), world.shapes.chain(pts, (f32(half + 0.5), pts[0][1]), (f32(-half - 0.5), pts[-1][1])))
    loop = [(f32(2.0 * math.cos(2 * math.pi * i / 12)), f32(9.0 + 1.0 * math.sin(2 * math.pi * i / 12))) for i in range(12)]
    ground.create_fixture(FixtureDef(friction=0.4), world.shapes.chain(loop, (0.0, 0.0), (0.0, 0.0), loop=True))
    shapes = [world.shapes.circle(0.35), world.shapes.polygon_box(0.4, 0.3), world.shapes.circle(0.2),
              world.shapes.polygon([(-0.4, -0.3), (0.4, -0.3), (0.0, 0.45)]),
              world.shapes.polygon([(f32(0.4 * math.cos(2 * math.pi * i / 8)), f32(0.4 * math.sin(2 * math.pi * i / 8))) for i in range(8)])]
    bodies = []
    for k in range(n):
        px = f32(-0.2 * segments + 0.4 * segments * ((k * 37) % n) / n + rng.uniform(-0.3, 0.3))
        py = f32(12.0 + 1.1 * (k % 7) + rng.uniform(-0.2, 0.2))
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py), angle=f32(rng.uniform(-3.0, 3.0)),
                                      linear_velocity=(f32(rng.uniform(-2.0, 2.0)), 0.0)))
        b.create_fixture(FixtureDef(density=1.0, friction=0.4, restitution=0.1 if k % 5 == 0 else 0.0), shapes[int(rng.next() % 5)])
        bodies.append(b)
    return bodies


# ------------------------------------------------------------------------------------------------------------
# scenes with joints: the reference has no joint test, so these follow its testbed recipes
# ------------------------------------------------------------------------------------------------------------
def bridge(world, count=30, testbed_ground_body=True):
    """examples/testbed/tests/bridge.rs:58-141: a plank bridge of `count` boxes chained by revolute joints between two
    anchors on the ground body, two triangles and three circles dropped on it."""
    if testbed_ground_body:
        world.create_body(BodyDef())
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
    plank = world.shapes.polygon_box(0.5, 0.125)
    prev = ground
    for i in range(count):
        body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-14.5 + 1.0 * i), 5.0)))
        body.create_fixture(FixtureDef(density=20.0, friction=0.2), plank)
        world.create_joint(world.revolute_joint_def(prev, body, (f32(-15.0 + 1.0 * i), 5.0)))
        prev = body
    world.create_joint(world.revolute_joint_def(prev, ground, (f32(-15.0 + 1.0 * count), 5.0)))
    tri = world.shapes.polygon([(-0.5, 0.0), (0.5, 0.0), (0.0, 1.5)])
    for i in range(2):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-8.0 + 8.0 * i), 12.0)))
        b.create_fixture(FixtureDef(density=1.0), tri)
    ball = world.shapes.circle(0.5)
    for i in range(3):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-6.0 + 6.0 * i), 10.0)))
        b.create_fixture(FixtureDef(density=1.0), ball)


def cantilever(world, count=8, testbed_ground_body=True):
    """examples/testbed/tests/cantilever.rs:60-243: four beams of planks chained by weld joints — rigid from the ground, soft
    (5 Hz) from the ground, rigid and soft (8 Hz) free chains — with two triangles and two circles dropped on them."""
    if testbed_ground_body:
        world.create_body(BodyDef())
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
    plank = world.shapes.polygon_box(0.5, 0.125)
    prev = ground
    for i in range(count):
        body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-14.5 + 1.0 * i), 5.0)))
        body.create_fixture(FixtureDef(density=20.0), plank)
        world.create_joint(world.weld_joint_def(prev, body, (f32(-15.0 + 1.0 * i), 5.0)))
        prev = body
    long_plank = world.shapes.polygon_box(1.0, 0.125)
    prev = ground
    for i in range(3):
        body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-14.0 + 2.0 * i), 15.0)))
        body.create_fixture(FixtureDef(density=20.0), long_plank)
        jd = world.weld_joint_def(prev, body, (f32(-15.0 + 2.0 * i), 15.0))
        jd.stiffness, jd.damping = world.angular_stiffness(5.0, 0.7, prev, body)
        world.create_joint(jd)
        prev = body
    prev = ground
    for i in range(count):
        body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-4.5 + 1.0 * i), 5.0)))
        body.create_fixture(FixtureDef(density=20.0), plank)
        if i > 0:
            world.create_joint(world.weld_joint_def(prev, body, (f32(-5.0 + 1.0 * i), 5.0)))
        prev = body
    prev = ground
    for i in range(count):
        body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(5.5 + 1.0 * i), 10.0)))
        body.create_fixture(FixtureDef(density=20.0), plank)
        if i > 0:
            jd = world.weld_joint_def(prev, body, (f32(5.0 + 1.0 * i), 10.0))
            jd.stiffness, jd.damping = world.angular_stiffness(8.0, 0.7, prev, body)
            world.create_joint(jd)
        prev = body
    tri = world.shapes.polygon([(-0.5, 0.0), (0.5, 0.0), (0.0, 1.5)])
    for i in range(2):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-8.0 + 8.0 * i), 12.0)))
        b.create_fixture(FixtureDef(density=1.0), tri)
    ball = world.shapes.circle(0.5)
    for i in range(2):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-6.0 + 6.0 * i), 10.0)))
        b.create_fixture(FixtureDef(density=1.0), ball)


def sliders(world):
    """Prismatic joints: examples/testbed/tests/prismatic_joint.rs:61-101 (a 2 x 2 box, turned a quarter, on a horizontal slider
    from the ground with limits +-10 and a motor that is off), plus a vertical lift with its motor on against a stack it
    carries, a slider between two dynamic bodies, and one with equal limits (a locked translation)."""
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
    body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 10.0), angle=f32(0.5 * math.pi), allow_sleep=0))
    body.create_fixture_by_shape(world.shapes.polygon_box(1.0, 1.0), 5.0)
    jd = world.prismatic_joint_def(ground, body, (0.0, 10.0), (1.0, 0.0))
    jd.motor_speed, jd.max_motor_torque, jd.enable_motor = 10.0, 10000.0, 0
    jd.lower_angle, jd.upper_angle, jd.enable_limit = -10.0, 10.0, 1
    world.create_joint(jd)
    body.set_linear_velocity((6.0, 0.0))
    # a lift: platform on a vertical slider, motor on, two boxes riding it
    lift = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-12.0, 2.0)))
    lift.create_fixture(FixtureDef(density=2.0, friction=0.6), world.shapes.polygon_box(2.0, 0.25))
    jd = world.prismatic_joint_def(ground, lift, (-12.0, 2.0), (0.0, 1.0))
    jd.motor_speed, jd.max_motor_torque, jd.enable_motor = 1.5, 2000.0, 1
    jd.lower_angle, jd.upper_angle, jd.enable_limit = 0.0, 6.0, 1
    world.create_joint(jd)
    for i in range(2):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(-12.5 + 1.0 * i), f32(2.76 + 1.02 * i))))
        b.create_fixture(FixtureDef(density=1.0, friction=0.6), world.shapes.polygon_box(0.5, 0.5))
    # a slider between two dynamic bodies, along a slanted axis
    a = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(10.0, 6.0), angle=0.3))
    a.create_fixture_by_shape(world.shapes.polygon_box(1.5, 0.3), 1.0)
    b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(12.0, 7.0)))
    b.create_fixture_by_shape(world.shapes.circle(0.5), 2.0)
    jd = world.prismatic_joint_def(a, b, (11.0, 6.5), (2.0, 1.0))
    jd.lower_angle, jd.upper_angle, jd.enable_limit = -0.5, 1.5, 1
    jd.collide_connected = 1
    world.create_joint(jd)
    # equal limits: the translation is locked
    c = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(20.0, 4.0)))
    c.create_fixture_by_shape(world.shapes.polygon_box(0.5, 0.5), 1.0)
    jd = world.prismatic_joint_def(ground, c, (20.0, 4.0), (0.0, 1.0))
    jd.lower_angle, jd.upper_angle, jd.enable_limit = 0.0, 0.0, 1
    world.create_joint(jd)


def car(world, motor_speed=-20.0):
    """The vehicle of examples/testbed/tests/car.rs:196-275 — a six-vertex chassis on two circle wheels hung on wheel joints
    along the chassis' y axis (4 Hz, damping ratio 0.7, limits +-0.25, rear wheel motorised with 20 Nm) — on a short track:
    flat ground, the testbed's first row of bumps (:88-101), a ramp and a stack of boxes to run into."""
    ground = world.create_body(BodyDef())
    gfd = FixtureDef(density=0.0, friction=0.6)
    ground.create_fixture(gfd, world.shapes.edge_two_sided((-20.0, 0.0), (20.0, 0.0)))
    hs = (0.25, 1.0, 4.0, 0.0, 0.0, -1.0, -2.0, -2.0, -1.25, 0.0)
    x, y1, dx = 20.0, 0.0, 5.0
    for i in range(10):
        y2 = hs[i]
        ground.create_fixture(gfd, world.shapes.edge_two_sided((f32(x), f32(y1)), (f32(x + dx), f32(y2))))
        y1 = y2
        x += dx
    ground.create_fixture(gfd, world.shapes.edge_two_sided((f32(x), 0.0), (f32(x + 40.0), 0.0)))
    x += 40.0
    ground.create_fixture(gfd, world.shapes.edge_two_sided((f32(x), 0.0), (f32(x + 10.0), 5.0)))
    box = world.shapes.polygon_box(0.5, 0.5)
    for i in range(4):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(12.0, f32(0.5 + 1.0 * i))))
        b.create_fixture(FixtureDef(density=0.5), box)
    chassis = world.shapes.polygon([(-1.5, -0.5), (1.5, -0.5), (1.5, 0.0), (0.0, 0.9), (-1.15, 0.9), (-1.5, 0.2)])
    circle = world.shapes.circle(0.4)
    car_body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 1.0)))
    car_body.create_fixture_by_shape(chassis, 1.0)
    wfd = FixtureDef(density=1.0, friction=0.9)
    wheel1 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-1.0, 0.35)))
    wheel1.create_fixture(wfd, circle)
    wheel2 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(1.0, 0.4)))
    wheel2.create_fixture(wfd, circle)
    hertz, ratio = f32(4.0), f32(0.7)
    omega = f32(f32(2.0) * f32(math.pi) * hertz)
    joints = []
    for wheel, pos, torque, motor in ((wheel1, (-1.0, 0.35), 20.0, 1), (wheel2, (1.0, 0.4), 10.0, 0)):
        mass = f32(f32(f32(f32(1.0) * f32(math.pi)) * f32(0.4)) * f32(0.4))  # B2circleShape::compute_mass: density * pi * r * r
        jd = world.wheel_joint_def(car_body, wheel, pos, (0.0, 1.0))
        jd.motor_speed, jd.max_motor_torque, jd.enable_motor = (motor_speed if motor else 0.0), torque, motor
        jd.stiffness = f32(f32(mass * omega) * omega)
        jd.damping = f32(f32(f32(f32(2.0) * mass) * ratio) * omega)
        jd.lower_angle, jd.upper_angle, jd.enable_limit = -0.25, 0.25, 1
        joints.append(world.create_joint(jd))
    return joints


def top_down(world):
    """Zero-gravity, top-down: examples/testbed/tests/apply_force.rs:106-140 (boxes held back by friction joints to the ground:
    max_force = m g, max_torque = 0.2 I g) kicked into each other, and examples/testbed/tests/motor_joint.rs:78-100 (a box driven
    to an offset pose by a motor joint: max_force 1000, max_torque 1000) with a target away from where it starts."""
    ground = world.create_body(BodyDef(position=(0.0, 20.0)))
    for a, b in (((-20.0, -20.0), (-20.0, 20.0)), ((20.0, -20.0), (20.0, 20.0)), ((-20.0, 20.0), (20.0, 20.0)), ((-20.0, -20.0), (20.0, -20.0))):
        ground.create_fixture(FixtureDef(density=0.0, restitution=0.4), world.shapes.edge_two_sided(a, b))
    box = world.shapes.polygon_box(0.5, 0.5)
    gravity = f32(10.0)
    for i in range(10):
        body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, f32(7.0 + 1.54 * i))))
        body.create_fixture(FixtureDef(density=1.0, friction=0.3), box)
        mass = f32(1.0)                         # 1 x 1 box of density 1
        inertia = f32(f32(1.0) * f32(f32(1.0 + 1.0) / f32(12.0)))  # m (w^2 + h^2) / 12
        radius = f32(math.sqrt(f32(f32(2.0) * inertia) / mass))
        jd = world.friction_joint_def(ground, body, (0.0, f32(7.0 + 1.54 * i)))
        jd.local_anchor_a[0], jd.local_anchor_a[1] = 0.0, 0.0
        jd.local_anchor_b[0], jd.local_anchor_b[1] = 0.0, 0.0
        jd.collide_connected = 1
        jd.length = f32(mass * gravity)                                   # max_force
        jd.max_motor_torque = f32(f32(f32(0.2) * f32(mass * radius)) * gravity)  # max_torque
        world.create_joint(jd)
        body.set_linear_velocity((f32(26.0 - 6.0 * i), f32(-18.0 + 4.5 * i)))  # fast: they slide and collide for seconds
        body.set_angular_velocity(f32(3.0 * i - 12.0))
    puck = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-8.0, 28.0)))
    puck.create_fixture(FixtureDef(density=2.0, friction=0.6), world.shapes.polygon_box(2.0, 0.5))
    jd = world.motor_joint_def(ground, puck)
    jd.local_anchor_a[0], jd.local_anchor_a[1] = 6.0, 4.0   # linear_offset: where body B shall go, in the ground body's frame
    jd.reference_angle = 1.0                                 # angular_offset
    jd.length, jd.max_motor_torque = 1000.0, 1000.0
    world.create_joint(jd)


def pulleys(world):
    """examples/testbed/tests/pulley_joint.rs:55-115 (two 1 x 2 boxes of density 5 on a ratio-1.5 pulley under two ground
    circles) next to a second, lighter pair over a floor so one side lands, and the testbed's mouse drag
    (examples/testbed/test.rs:230-262: 5 Hz, damping ratio 0.7, max_force 1000 m) pulling a box sideways off the floor.
    Returns the mouse joint so a test can move its target."""
    y, l, a, b = 16.0, 12.0, 1.0, 2.0
    ground = world.create_body(BodyDef())
    for x in (-10.0, 10.0):
        ground.create_fixture(FixtureDef(density=0.0), world.shapes.circle(2.0, (x, y + b + l)))
    ground.create_fixture(FixtureDef(density=0.0), world.shapes.edge_two_sided((-60.0, 0.0), (60.0, 0.0)))
    shape = world.shapes.polygon_box(a, b)
    body1 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-10.0, y)))
    body1.create_fixture_by_shape(shape, 5.0)
    body2 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(10.0, y)))
    body2.create_fixture_by_shape(shape, 5.0)
    world.create_joint(world.pulley_joint_def(body1, body2, (-10.0, y + b + l), (10.0, y + b + l), (-10.0, y + b), (10.0, y + b), 1.5))
    # a lighter pair with a block-and-tackle ratio: the heavy side comes down on the floor, the rope goes slack-free
    small, big = world.shapes.polygon_box(0.5, 0.5), world.shapes.polygon_box(1.0, 1.0)
    body3 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(24.0, 6.0), angle=0.2))
    body3.create_fixture(FixtureDef(density=1.0, friction=0.4), small)
    body4 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(32.0, 9.0)))
    body4.create_fixture(FixtureDef(density=2.0, friction=0.4), big)
    world.create_joint(world.pulley_joint_def(body3, body4, (25.0, 20.0), (31.0, 20.0), (24.0, 6.5), (32.0, 10.0), 2.0))
    # mouse drag
    crate = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-30.0, 1.0)))
    crate.create_fixture(FixtureDef(density=1.0, friction=0.5), big)
    jd = world.mouse_joint_def(ground, crate, (-29.5, 1.5))
    jd.length = f32(1000.0 * 4.0)  # max_force = 1000 * mass (2 x 2 box of density 1)
    jd.stiffness, jd.damping = world.linear_stiffness(5.0, 0.7, ground, crate)
    mouse = world.create_joint(jd)
    mouse.set_target((-22.0, 9.0))
    crate.set_awake(True)
    return mouse


def gears(world):
    """examples/testbed/tests/gear_joint.rs:66-205: a pendulum bar hinged on a static disc with a disc hinged on its end, the two
    hinges geared 2 : 1 (the testbed hands the gear joint the STATIC disc as body A — kept: the port solves on the def's
    bodies), and the ground-mounted train disc - disc - rack (two revolute joints and a prismatic joint with limits, gear
    ratios 2 and -1/2).  A kick on the small disc and on the bar sets everything turning.  Returns the first train gear."""
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((50.0, 0.0), (-50.0, 0.0)), 0.0)
    circle1, circle2, box = world.shapes.circle(1.0), world.shapes.circle(2.0), world.shapes.polygon_box(0.5, 5.0)
    body1 = world.create_body(BodyDef(type=abi.STATIC_BODY, position=(10.0, 9.0)))
    body1.create_fixture_by_shape(circle1, 5.0)
    body2 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(10.0, 8.0)))
    body2.create_fixture_by_shape(box, 5.0)
    body3 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(10.0, 6.0)))
    body3.create_fixture_by_shape(circle2, 5.0)
    joint1 = world.create_joint(world.revolute_joint_def(body1, body2, (10.0, 9.0)))
    joint2 = world.create_joint(world.revolute_joint_def(body2, body3, (10.0, 6.0)))
    jd = world.gear_joint_def(joint1, joint2, 2.0)
    jd.body_a, jd.body_b = body1.index, body3.index
    world.create_joint(jd)
    body2.set_angular_velocity(1.5)
    # the train
    b1 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-3.0, 12.0)))
    b1.create_fixture_by_shape(circle1, 5.0)
    j1 = world.create_joint(world.revolute_joint_def(ground, b1, (-3.0, 12.0)))
    b2 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 12.0)))
    b2.create_fixture_by_shape(circle2, 5.0)
    j2 = world.create_joint(world.revolute_joint_def(ground, b2, (0.0, 12.0)))
    b3 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(2.5, 12.0)))
    b3.create_fixture_by_shape(box, 5.0)
    jd3 = world.prismatic_joint_def(ground, b3, (2.5, 12.0), (0.0, 1.0))
    jd3.lower_angle, jd3.upper_angle, jd3.enable_limit = -5.0, 5.0, 1
    j3 = world.create_joint(jd3)
    train = world.create_joint(world.gear_joint_def(j1, j2, 2.0))
    world.create_joint(world.gear_joint_def(j2, j3, -0.5))
    b1.set_angular_velocity(6.0)
    return train


def tumbler(world, n=200, seed=0xB2D + 21):
    """examples/testbed/tests/tumbler.rs:62-97: a hollow box of four plank fixtures turned by a revolute-joint motor
    (0.05 pi rad/s, torque 1e8) around a point of the ground body; the testbed drops one 0.125 box per step, here `n`
    boxes start inside on a jittered grid (the step itself stays deterministic)."""
    rng = SplitMix64(seed)
    ground = world.create_body(BodyDef())
    body = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 10.0), allow_sleep=0))
    for hx, hy, cx, cy in ((0.5, 10.0, 10.0, 0.0), (0.5, 10.0, -10.0, 0.0), (10.0, 0.5, 0.0, 10.0), (10.0, 0.5, 0.0, -10.0)):
        body.create_fixture_by_shape(world.shapes.polygon_box(hx, hy, center=(cx, cy), angle=0.0), 5.0)
    jd = world.revolute_joint_def(ground, body, (0.0, 10.0))
    jd.local_anchor_a[0], jd.local_anchor_a[1] = 0.0, 10.0
    jd.local_anchor_b[0], jd.local_anchor_b[1] = 0.0, 0.0
    jd.reference_angle = 0.0
    jd.motor_speed = f32(np.float32(0.05) * np.float32(math.pi))
    jd.max_motor_torque = 1e8
    jd.enable_motor = 1
    joint = world.create_joint(jd)
    small = world.shapes.polygon_box(0.125, 0.125)
    cols = 20
    for k in range(n):
        px = f32(-4.0 + 0.4 * (k % cols) + rng.uniform(-0.05, 0.05))
        py = f32(3.0 + 0.4 * (k // cols) + rng.uniform(-0.05, 0.05))
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py)))
        b.create_fixture_by_shape(small, 1.0)
    return joint


def joints_mix(world, seed=0xB2D + 23):
    """Every joint branch of the step in one small scene: a rigid distance-joint pendulum, a soft distance joint
    (b2_linear_stiffness, min < max: spring + lower + upper rows), a revolute chain with limits hanging from a static
    anchor (examples/testbed/tests/chain.rs), a limited + motorised revolute arm, a joint with collide_connected, a
    fixed-rotation body on a revolute joint, and loose boxes that collide with all of it."""
    rng = SplitMix64(seed)
    ground = world.create_body(BodyDef())
    ground.create_fixture_by_shape(world.shapes.edge_two_sided((-30.0, 0.0), (30.0, 0.0)), 0.0)
    ball = world.shapes.circle(0.4)
    # rigid pendulum (equal limits)
    p1 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-12.0, 8.0)))
    p1.create_fixture(FixtureDef(density=1.0, friction=0.3), ball)
    world.create_joint(world.distance_joint_def(ground, p1, (-15.0, 12.0), (-12.0, 8.0)))
    # soft spring with slack between min and max length
    p2 = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-8.0, 9.0), angular_damping=0.1))
    p2.create_fixture(FixtureDef(density=2.0, friction=0.3), world.shapes.polygon_box(0.5, 0.3))
    jd = world.distance_joint_def(ground, p2, (-8.0, 13.0), (-8.0, 9.5))
    jd.stiffness, jd.damping = world.linear_stiffness(2.0, 0.3, ground, p2)
    jd.min_length = f32(jd.length - 1.0)
    jd.max_length = f32(jd.length + 0.5)
    world.create_joint(jd)
    # revolute chain with limits (chain.rs: friction 0.2, density 20; here every joint is limited to +-0.6 rad)
    link = world.shapes.polygon_box(0.6, 0.125)
    prev = ground
    y = 14.0
    for i in range(12):
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(f32(0.5 + i * 1.0), y)))
        b.create_fixture(FixtureDef(density=20.0, friction=0.2), link)
        jd = world.revolute_joint_def(prev, b, (f32(i * 1.0), y))
        if i > 0:
            jd.enable_limit = 1
            jd.lower_angle = -0.6
            jd.upper_angle = 0.6
        world.create_joint(jd)
        prev = b
    # motorised, limited arm; its two bodies may collide (collide_connected)
    base = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-3.0, 1.0)))
    base.create_fixture(FixtureDef(density=5.0, friction=0.6), world.shapes.polygon_box(1.0, 1.0))
    arm = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-3.0, 3.5)))
    arm.create_fixture(FixtureDef(density=1.0, friction=0.3), world.shapes.polygon_box(0.2, 1.5))
    jd = world.revolute_joint_def(base, arm, (-3.0, 2.0))
    jd.collide_connected = 1
    jd.enable_limit, jd.lower_angle, jd.upper_angle = 1, -1.0, 0.8
    jd.enable_motor, jd.motor_speed, jd.max_motor_torque = 1, 1.5, 40.0
    motor = world.create_joint(jd)
    # fixed-rotation body on a revolute joint with a very narrow limit range (the "equal limits" correction branch)
    fr = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(6.0, 5.0), fixed_rotation=1))
    fr.create_fixture(FixtureDef(density=1.0), world.shapes.polygon_box(0.4, 0.4))
    nb = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(7.5, 5.0)))
    nb.create_fixture(FixtureDef(density=1.0), world.shapes.polygon_box(0.9, 0.2))
    jd = world.revolute_joint_def(fr, nb, (6.6, 5.0))
    jd.enable_limit, jd.lower_angle, jd.upper_angle = 1, -0.01, 0.01
    world.create_joint(jd)
    world.create_joint(world.distance_joint_def(ground, fr, (6.0, 9.0), (6.0, 5.0)))
    # loose bodies
    box = world.shapes.polygon_box(0.35, 0.35)
    for k in range(24):
        px = f32(-10.0 + 0.9 * k + rng.uniform(-0.1, 0.1))
        py = f32(16.0 + (k % 3) * 1.1 + rng.uniform(-0.1, 0.1))
        b = world.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, py), angle=f32(rng.uniform(-1.0, 1.0))))
        b.create_fixture(FixtureDef(density=1.0, friction=0.4), box if k % 2 else ball)
    return motor
