"""Joints inside the island solve: revolute and distance joints.

The reference has NO joint test (tests/ only covers a falling box, begin_contact, polygon mass and sweeps), so the
oracle's joint code is pinned by analytic known answers here and is otherwise a literal restatement of
src/private/dynamics/joints/b2_{revolute,distance}_joint.rs ("parity unpinned" applies to it as to the rest).  The
device stages (host simulator here, CUDA in the tests marked gpu) must equal the oracle bit for bit."""
import math

import numpy as np
import pytest

import parity
from conftest import HOSTSIM_SO, JOINT_SCENES


@pytest.fixture(scope="module")
def hctx(built):
    from box2d_rs_b200 import batch
    c = batch.Context(0, lib_path=HOSTSIM_SO)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gctx(built):
    from box2d_rs_b200 import batch
    c = batch.Context(0)
    yield c
    c.close()


def _pair(name, ctx):
    from box2d_rs_b200 import scenes, world
    from oracle import b2o
    recipe, gravity, steps = JOINT_SCENES[name]
    wo = b2o.B2world(gravity)
    ro = recipe(scenes, wo)
    wg = world.B2world(gravity, ctx=ctx)
    rg = recipe(scenes, wg)
    return wo, wg, steps, ro, rg


# ---------------------------------------------------------------------------------------------- oracle known answers
def test_oracle_distance_joint_keeps_its_length(built):
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    bob = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(3.0, 5.0)))
    bob.create_fixture_by_shape(w.shapes.circle(0.5), 1.0)
    w.create_joint(w.distance_joint_def(ground, bob, (0.0, 5.0), (3.0, 5.0)))
    lowest = 5.0
    for _ in range(240):
        w.step(scenes.DT, 8, 3)
        c = w.snapshot().bodies[1]["c"]
        assert abs(math.hypot(c[0], c[1] - 5.0) - 3.0) < 0.005  # within the linear slop
        lowest = min(lowest, float(c[1]))
    assert lowest < 2.2  # it really swung through the bottom (y = 2)


def test_oracle_revolute_joint_pins_the_anchor_and_respects_limits(built):
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    arm = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(1.0, 4.0)))
    arm.create_fixture_by_shape(w.shapes.polygon_box(1.0, 0.1), 1.0)
    jd = w.revolute_joint_def(ground, arm, (0.0, 4.0))
    jd.enable_limit, jd.lower_angle, jd.upper_angle = 1, -0.5, 0.25
    w.create_joint(jd)
    for _ in range(180):
        w.step(scenes.DT, 8, 3)
        b = w.snapshot().bodies[1]
        # anchor of the arm (local (-1, 0)) stays on the ground anchor (0, 4)
        ax = b["xf"][0] + b["xf"][3] * -1.0
        ay = b["xf"][1] + b["xf"][2] * -1.0
        assert math.hypot(ax, ay - 4.0) < 0.01
        assert -0.5 - 0.05 <= b["a"] <= 0.25 + 0.05
    assert abs(w.snapshot().bodies[1]["a"] + 0.5) < 0.04  # resting on the lower limit


def test_oracle_revolute_motor_reaches_its_speed(built):
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, 0.0))
    ground = w.create_body(BodyDef())
    wheel = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.0)))
    wheel.create_fixture_by_shape(w.shapes.circle(1.0), 1.0)
    jd = w.revolute_joint_def(ground, wheel, (0.0, 0.0))
    jd.enable_motor, jd.motor_speed, jd.max_motor_torque = 1, 2.0, 1000.0
    j = w.create_joint(jd)
    for _ in range(30):
        w.step(scenes.DT, 8, 3)
    assert abs(w.snapshot().bodies[1]["w"] - 2.0) < 1e-4
    j.set_motor_speed(-1.0)
    for _ in range(30):
        w.step(scenes.DT, 8, 3)
    assert abs(w.snapshot().bodies[1]["w"] + 1.0) < 1e-4


def test_oracle_weld_joint_keeps_the_relative_pose(built):
    """Two free bodies welded together and thrown: the anchor points stay together and the relative angle stays the
    reference angle (b2_weld_joint.rs: rigid form, 3x3 solve); the soft form lets the angle oscillate and damps it out."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    for soft in (False, True):
        w = b2o.B2world((0.0, -10.0))
        a = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 10.0), angle=0.3))
        a.create_fixture_by_shape(w.shapes.polygon_box(1.0, 0.2), 2.0)
        b = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(2.0, 10.5), angle=-0.4))
        b.create_fixture_by_shape(w.shapes.polygon_box(0.5, 0.5), 1.0)
        jd = w.weld_joint_def(a, b, (1.0, 10.3))
        assert abs(jd.reference_angle - (-0.7)) < 1e-6
        if soft:
            jd.stiffness, jd.damping = w.angular_stiffness(3.0, 0.5, a, b)
        w.create_joint(jd)
        a.set_angular_velocity(2.0)
        b.apply_linear_impulse_to_center((3.0, 4.0), True)
        worst_angle = 0.0
        for i in range(150):
            w.step(scenes.DT, 8, 3)
            ba, bb = w.snapshot().bodies[0], w.snapshot().bodies[1]

            def world_point(body, lp):
                return (body["xf"][0] + body["xf"][3] * lp[0] - body["xf"][2] * lp[1],
                        body["xf"][1] + body["xf"][2] * lp[0] + body["xf"][3] * lp[1])
            pa, pb = world_point(ba, jd.local_anchor_a), world_point(bb, jd.local_anchor_b)
            assert math.hypot(pa[0] - pb[0], pa[1] - pb[1]) < 0.02
            err = abs(float(bb["a"] - ba["a"]) - jd.reference_angle)
            if i > 100:
                worst_angle = max(worst_angle, err)
            if not soft:
                assert err < 0.02
        assert worst_angle < (0.05 if soft else 0.01)


def test_oracle_prismatic_joint_slides_along_its_axis_only(built):
    """examples/testbed/tests/prismatic_joint.rs: a box on a horizontal slider from the ground, gravity on, pushed along the
    axis — it keeps its height and its angle, runs into the upper limit (10) and rests there; with the motor on it is driven
    back at the motor speed."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    box = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 10.0), angle=0.5 * math.pi, allow_sleep=0))
    box.create_fixture_by_shape(w.shapes.polygon_box(1.0, 1.0), 5.0)
    jd = w.prismatic_joint_def(ground, box, (0.0, 10.0), (1.0, 0.0))
    assert abs(jd.length - 1.0) < 1e-7 and abs(jd.min_length) < 1e-7  # local axis of the (unrotated) ground body
    jd.lower_angle, jd.upper_angle, jd.enable_limit = -10.0, 10.0, 1
    jd.motor_speed, jd.max_motor_torque = -2.0, 10000.0
    j = w.create_joint(jd)
    box.set_linear_velocity((8.0, 0.0))
    far = 0.0
    for _ in range(240):
        w.step(scenes.DT, 8, 3)
        b = w.snapshot().bodies[1]
        assert abs(b["c"][1] - 10.0) < 0.01 and abs(b["a"] - 0.5 * math.pi) < 0.01
        assert b["c"][0] < 10.0 + 0.02
        far = max(far, float(b["c"][0]))
    assert far > 9.9  # it reached the limit
    j.enable_motor(True)
    for _ in range(60):
        w.step(scenes.DT, 8, 3)
    b = w.snapshot().bodies[1]
    assert abs(b["v"][0] + 2.0) < 1e-3 and abs(b["v"][1]) < 1e-3


def test_oracle_prismatic_motor_lifts_against_gravity_until_the_limit(built):
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    lift = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 2.0), allow_sleep=0))
    lift.create_fixture_by_shape(w.shapes.polygon_box(2.0, 0.25), 2.0)  # mass 4: weight 40 < max motor force
    jd = w.prismatic_joint_def(ground, lift, (0.0, 2.0), (0.0, 1.0))
    jd.motor_speed, jd.max_motor_torque, jd.enable_motor = 1.5, 2000.0, 1
    jd.lower_angle, jd.upper_angle, jd.enable_limit = 0.0, 3.0, 1
    w.create_joint(jd)
    for i in range(60):
        w.step(scenes.DT, 8, 3)
    b = w.snapshot().bodies[1]
    assert abs(b["v"][1] - 1.5) < 1e-3 and abs(b["c"][0]) < 1e-3  # rising at the motor speed, on the axis
    for i in range(120):
        w.step(scenes.DT, 8, 3)
    b = w.snapshot().bodies[1]
    assert abs(b["c"][1] - 5.0) < 0.02 and abs(b["v"][1]) < 1e-2  # held at the upper limit (2 + 3)


def test_oracle_wheel_joints_carry_a_car(built):
    """The vehicle of examples/testbed/tests/car.rs on flat ground: the rear wheel's motor (-20 rad/s, radius 0.4) brings the
    car to 8 m/s; each wheel stays on its chassis axis (point-to-line constraint) within its translation limits."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef, FixtureDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    ground.create_fixture(FixtureDef(density=0.0, friction=0.6), w.shapes.edge_two_sided((-20.0, 0.0), (400.0, 0.0)))
    chassis = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 1.0)))
    chassis.create_fixture_by_shape(w.shapes.polygon([(-1.5, -0.5), (1.5, -0.5), (1.5, 0.0), (0.0, 0.9), (-1.15, 0.9), (-1.5, 0.2)]), 1.0)
    wheels, defs = [], []
    for pos, torque, motor in (((-1.0, 0.35), 20.0, 1), ((1.0, 0.4), 10.0, 0)):
        wh = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=pos))
        wh.create_fixture(FixtureDef(density=1.0, friction=0.9), w.shapes.circle(0.4))
        jd = w.wheel_joint_def(chassis, wh, pos, (0.0, 1.0))
        assert abs(jd.length) < 1e-7 and abs(jd.min_length - 1.0) < 1e-7  # local axis = (0, 1): the chassis is not rotated
        jd.motor_speed, jd.max_motor_torque, jd.enable_motor = (-20.0 if motor else 0.0), torque, motor
        jd.stiffness, jd.damping = w.linear_stiffness(4.0, 0.7, chassis, wh)
        jd.lower_angle, jd.upper_angle, jd.enable_limit = -0.25, 0.25, 1
        w.create_joint(jd)
        wheels.append(wh)
        defs.append(jd)
    for i in range(300):
        w.step(scenes.DT, 8, 3)
        if i % 20 == 19:
            b = w.snapshot().bodies
            c = b[1]
            for k, jd in enumerate(defs):
                wh = b[2 + k]
                # anchor on the chassis and the chassis' y axis in world space
                ax = c["xf"][0] + c["xf"][3] * jd.local_anchor_a[0] - c["xf"][2] * jd.local_anchor_a[1]
                ay = c["xf"][1] + c["xf"][2] * jd.local_anchor_a[0] + c["xf"][3] * jd.local_anchor_a[1]
                ux, uy = -c["xf"][2], c["xf"][3]  # rotated (0, 1)
                dx, dy = wh["xf"][0] - ax, wh["xf"][1] - ay
                assert abs(dx * uy - dy * ux) < 0.02              # on the line
                assert -0.25 - 0.02 <= dx * ux + dy * uy <= 0.25 + 0.02  # inside the limits
    b = w.snapshot().bodies
    assert abs(b[1]["v"][0] - 8.0) < 0.4 and abs(b[2]["w"] + 20.0) < 0.5 and b[1]["c"][0] > 25.0


def test_oracle_friction_joint_decelerates_at_max_force_over_mass(built):
    """examples/testbed/tests/apply_force.rs: top-down friction, no gravity.  A sliding box loses max_force / m per second until
    it stops (Coulomb friction in the plane), its spin loses max_torque / I per second."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, 0.0))
    ground = w.create_body(BodyDef())
    box = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.0), allow_sleep=0))
    box.create_fixture_by_shape(w.shapes.polygon_box(0.5, 0.5), 2.0)  # mass 2, I = 2 / 6
    jd = w.friction_joint_def(ground, box, (0.0, 0.0))
    jd.length, jd.max_motor_torque = 4.0, 0.5  # max_force, max_torque
    w.create_joint(jd)
    box.set_linear_velocity((3.0, 4.0))  # speed 5
    box.set_angular_velocity(6.0)
    for i in range(60):
        w.step(scenes.DT, 8, 3)
    b = w.snapshot().bodies[1]
    speed = math.hypot(b["v"][0], b["v"][1])
    assert abs(speed - (5.0 - 4.0 / 2.0 * 1.0)) < 0.02       # a = F / m = 2
    assert abs(b["v"][0] / b["v"][1] - 0.75) < 1e-3          # straight line
    assert abs(b["w"] - (6.0 - 0.5 / (2.0 / 6.0) * 1.0)) < 0.02  # alpha = T / I = 1.5
    for i in range(120):
        w.step(scenes.DT, 8, 3)
    b = w.snapshot().bodies[1]
    assert math.hypot(b["v"][0], b["v"][1]) < 1e-4           # stopped after 2.5 s


def test_oracle_motor_joint_drives_to_its_offsets(built):
    """examples/testbed/tests/motor_joint.rs: body B is driven to linear_offset / angular_offset in body A's frame."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef(position=(1.0, 2.0)))
    box = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(3.0, 8.0), allow_sleep=0))
    box.create_fixture_by_shape(w.shapes.polygon_box(2.0, 0.5), 2.0)  # mass 8: weight 80 < max_force
    jd = w.motor_joint_def(ground, box)
    assert abs(jd.local_anchor_a[0] - 2.0) < 1e-6 and abs(jd.local_anchor_a[1] - 6.0) < 1e-6  # body B's position in A's frame
    assert abs(jd.length - 1.0) < 1e-7 and abs(jd.max_motor_torque - 1.0) < 1e-7 and abs(jd.stiffness - 0.3) < 1e-7  # defaults
    jd.local_anchor_a[0], jd.local_anchor_a[1], jd.reference_angle = 5.0, 4.0, 0.8
    jd.length, jd.max_motor_torque = 1000.0, 1000.0
    w.create_joint(jd)
    for i in range(240):
        w.step(scenes.DT, 8, 3)
    b = w.snapshot().bodies[1]
    # the spring-like correction (factor 0.3) holds the box a little below the target: 8 kg x 10 m/s^2 against it
    assert abs(b["c"][0] - 6.0) < 0.01 and 5.9 < b["c"][1] < 6.0 + 1e-3 and abs(b["a"] - 0.8) < 0.01
    assert math.hypot(b["v"][0], b["v"][1]) < 1e-3


def test_oracle_pulley_joint_keeps_length_a_plus_ratio_length_b(built):
    """examples/testbed/tests/pulley_joint.rs: equal boxes on a ratio-1.5 pulley.  length_a + ratio * length_b stays at its
    initial value (12 + 1.5 * 12 = 30) within the position solver's slop while the side with the mechanical advantage (B: its
    rope carries ratio x the tension) rises."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    shape = w.shapes.polygon_box(1.0, 2.0)
    b1 = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-10.0, 16.0), allow_sleep=0))
    b1.create_fixture_by_shape(shape, 5.0)
    b2 = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(10.0, 16.0), allow_sleep=0))
    b2.create_fixture_by_shape(shape, 5.0)
    jd = w.pulley_joint_def(b1, b2, (-10.0, 30.0), (10.0, 30.0), (-10.0, 18.0), (10.0, 18.0), 1.5)
    assert jd.collide_connected == 1 and abs(jd.length - 12.0) < 1e-6 and abs(jd.min_length - 12.0) < 1e-6 and jd.max_length == 1.5
    assert (jd.lower_angle, jd.upper_angle, jd.max_motor_torque, jd.motor_speed) == (-10.0, 30.0, 10.0, 30.0)  # ground anchors
    w.create_joint(jd)
    for i in range(60):
        w.step(scenes.DT, 8, 3)
        ba, bb = w.snapshot().bodies[1], w.snapshot().bodies[2]
        la = math.hypot(ba["c"][0] + 10.0, ba["c"][1] + 2.0 - 30.0)  # anchors stay above the centres: the boxes do not turn
        lb = math.hypot(bb["c"][0] - 10.0, bb["c"][1] + 2.0 - 30.0)
        assert abs(la + 1.5 * lb - 30.0) < 0.02
    assert ba["c"][1] < 15.0 and bb["c"][1] > 16.5 and abs(ba["a"]) < 1e-4
    # accelerations: T from m a1 = m g - T, m a2 = 1.5 T - m g, a1 = 1.5 a2  ->  a2 = g / 6.5: after 1 s body B rose ~0.77
    assert abs((bb["c"][1] - 16.0) - 0.5 * 10.0 / 6.5) < 0.05
    assert w.snapshot().joints[0]["impulse"][0] > 0.0  # rope in tension


def test_oracle_mouse_joint_drags_the_body_to_its_target(built):
    """examples/testbed/test.rs:230-262 (mouse_down / mouse_move): a soft constraint pulls the grabbed point to the target;
    set_target wakes body B; the force never exceeds max_force."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    ground = w.create_body(BodyDef())
    ground.create_fixture_by_shape(w.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
    box = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.5)))
    box.create_fixture_by_shape(w.shapes.polygon_box(0.5, 0.5), 1.0)
    for i in range(120):
        w.step(scenes.DT, 8, 3)
    assert not w.snapshot().bodies[1]["flags"] & abi.BODY_AWAKE  # asleep on the floor
    jd = w.mouse_joint_def(ground, box, (0.0, 0.5))
    assert (jd.local_anchor_a[0], jd.local_anchor_a[1], jd.length) == (0.0, 0.5, 0.0)
    jd.length = 1000.0  # max_force = 1000 * mass
    jd.stiffness, jd.damping = w.linear_stiffness(5.0, 0.7, ground, box)
    mouse = w.create_joint(jd)
    rec = w.snapshot().joints[0]
    assert abs(rec["local_anchor_b"][0]) < 1e-6 and abs(rec["local_anchor_b"][1]) < 0.02  # the grabbed point, body-local
    assert not w.snapshot().bodies[1]["flags"] & abi.BODY_AWAKE  # creating the joint doesn't wake
    mouse.set_target((0.0, 0.5))                                 # unchanged target: still asleep
    assert not w.snapshot().bodies[1]["flags"] & abi.BODY_AWAKE
    mouse.set_target((3.0, 4.0))
    assert w.snapshot().bodies[1]["flags"] & abi.BODY_AWAKE
    peak = 0.0
    for i in range(180):
        w.step(scenes.DT, 8, 3)
        imp = w.snapshot().joints[0]["impulse"]
        peak = max(peak, math.hypot(imp[0], imp[1]) * 60.0)
    b = w.snapshot().bodies[1]
    assert abs(b["c"][0] - 3.0) < 0.02 and 3.9 < b["c"][1] < 4.0 + 0.01  # hangs a little below: m g / k
    assert peak <= 1000.0 + 1e-3
    jd.length = 5.0  # weaker than the weight (10): cannot lift the box
    w2 = b2o.B2world((0.0, -10.0))
    g2 = w2.create_body(BodyDef())
    g2.create_fixture_by_shape(w2.shapes.edge_two_sided((-40.0, 0.0), (40.0, 0.0)), 0.0)
    box2 = w2.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.5)))
    box2.create_fixture_by_shape(w2.shapes.polygon_box(0.5, 0.5), 1.0)
    w2.create_joint(jd).set_target((0.0, 6.0))
    for i in range(120):
        w2.step(scenes.DT, 8, 3)
        imp = w2.snapshot().joints[0]["impulse"]
        assert math.hypot(imp[0], imp[1]) * 60.0 <= 5.0 + 1e-4
    assert w2.snapshot().bodies[1]["c"][1] < 0.6


def test_oracle_gear_joint_keeps_coordinate1_plus_ratio_coordinate2(built):
    """examples/testbed/tests/gear_joint.rs (the ground-mounted train): disc 1 and disc 2 on revolute joints geared 2 : 1, disc 2
    and a rack on a prismatic joint geared -1/2.  angle1 + 2 angle2 and angle2 - translation / 2 stay at their initial
    values (0) while the train turns, until the rack reaches its limit and stops everything."""
    from box2d_rs_b200 import abi, scenes
    from oracle import b2o
    w = b2o.B2world((0.0, -10.0))
    train = scenes.gears(w)
    jd = w.gear_joint_def(w.joint(3), w.joint(4), 2.0)
    assert (jd.type, jd.body_a, jd.body_b, jd.enable_limit, jd.enable_motor, jd.length) == (abi.JOINT_GEAR, 4, 5, 3, 4, 2.0)
    rec = w.snapshot().joints[train.index]
    assert rec["type"] == abi.JOINT_GEAR and rec["flags"] & 0x300 == 0 and rec["impulse"][4] == 2.0 and rec["impulse"][3] == 0.0
    assert np.array(rec["impulse"][5:7], np.float32).view(np.int32).tolist() == [0, 0]  # bodies C, D: the ground
    rack = w.snapshot().joints[train.index + 1]
    assert rack["flags"] & 0x300 == 0x200 and rack["impulse"][4] == -0.5 and tuple(rack["param"][6:8]) == (0.0, 1.0)
    turned = 0.0
    for i in range(120):
        w.step(scenes.DT, 8, 3)
        b = w.snapshot().bodies
        a1, a2, y3 = float(b[4]["a"]), float(b[5]["a"]), float(b[6]["c"][1])
        assert abs(a1 + 2.0 * a2) < 0.02 and abs(a2 - 0.5 * (y3 - 12.0)) < 0.02
        turned = max(turned, abs(a1))
    assert turned > 4.9 and abs(y3 - 7.0) < 0.02  # the rack came down to its lower limit (-5) and holds the train
    assert abs(float(b[4]["w"])) < 1e-3


def test_oracle_angular_stiffness_formula(built):
    """b2_angular_stiffness (private b2_joint.rs:47-70): I = Ia Ib / (Ia + Ib) of B2body::get_inertia, omega = 2 pi f."""
    from box2d_rs_b200 import abi
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    w = b2o.B2world((0.0, 0.0))
    ground = w.create_body(BodyDef())
    a = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.0)))
    a.create_fixture_by_shape(w.shapes.polygon_box(1.0, 0.125), 20.0)   # mass 10, I = 10 (4 + 1/16) / 12
    b = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(3.0, 0.0)))
    b.create_fixture_by_shape(w.shapes.circle(0.5), 4.0)                # mass pi, I = pi / 8
    ia, ib = 10.0 * (4.0 + 0.0625) / 12.0, math.pi * 0.125
    k, d = w.angular_stiffness(5.0, 0.7, ground, a)
    omega = 2.0 * math.pi * 5.0
    assert abs(k - ia * omega * omega) / k < 1e-5 and abs(d - 2.0 * ia * 0.7 * omega) / d < 1e-5
    k, d = w.angular_stiffness(2.0, 0.3, a, b)
    i = ia * ib / (ia + ib)
    omega = 2.0 * math.pi * 2.0
    assert abs(k - i * omega * omega) / k < 1e-5 and abs(d - 2.0 * i * 0.3 * omega) / d < 1e-5


def test_weld_defs_and_unsupported_types(built):
    """b2gpu_weld_joint_def / b2gpu_angular_stiffness equal the oracle's bit for bit; a joint type outside the supported set
    is refused with B2GPU_E_UNSUPPORTED."""
    from box2d_rs_b200 import abi, batch, world
    from box2d_rs_b200.abi import BodyDef
    from box2d_rs_b200.lib import B2gpuError
    from oracle import b2o
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    ws = []
    for mk in (lambda: b2o.B2world((0.0, -10.0)), lambda: world.B2world((0.0, -10.0), ctx=ctx)):
        w = mk()
        a = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.25, 1.5), angle=0.7))
        a.create_fixture_by_shape(w.shapes.polygon_box(1.0, 0.125, center=(0.3, 0.1), angle=0.2), 20.0)
        b = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(2.0, 1.0), angle=-1.1))
        b.create_fixture_by_shape(w.shapes.circle(0.5), 3.0)
        jd = w.weld_joint_def(a, b, (1.3, 1.2))
        ws.append((w, jd, w.angular_stiffness(4.0, 0.6, a, b)))
    (wo, jo, so), (wg, jg, sg) = ws
    assert bytes(jo) == bytes(jg)
    assert np.array_equal(np.float32(so).view(np.uint32), np.float32(sg).view(np.uint32))
    po = wo.prismatic_joint_def(wo.body(0), wo.body(1), (1.3, 1.2), (0.6, -0.8))
    pg = wg.prismatic_joint_def(wg.body(0), wg.body(1), (1.3, 1.2), (0.6, -0.8))
    assert bytes(po) == bytes(pg)
    qo = wo.wheel_joint_def(wo.body(0), wo.body(1), (1.3, 1.2), (0.6, -0.8))
    qg = wg.wheel_joint_def(wg.body(0), wg.body(1), (1.3, 1.2), (0.6, -0.8))
    assert bytes(qo) == bytes(qg)
    assert bytes(wo.friction_joint_def(wo.body(0), wo.body(1), (1.3, 1.2))) == bytes(wg.friction_joint_def(wg.body(0), wg.body(1), (1.3, 1.2)))
    assert bytes(wo.motor_joint_def(wo.body(0), wo.body(1))) == bytes(wg.motor_joint_def(wg.body(0), wg.body(1)))
    args = (wo.body(0), wo.body(1), (0.5, 6.0), (2.5, 7.0), (0.4, 1.7), (2.1, 1.2), 2.5)
    argsg = (wg.body(0), wg.body(1)) + args[2:]
    assert bytes(wo.pulley_joint_def(*args)) == bytes(wg.pulley_joint_def(*argsg))
    assert bytes(wo.mouse_joint_def(wo.body(0), wo.body(1), (2.2, 1.1))) == bytes(wg.mouse_joint_def(wg.body(0), wg.body(1), (2.2, 1.1)))
    with pytest.raises(B2gpuError) as e:  # ratio <= epsilon: the reference asserts
        wg.pulley_joint_def(*(argsg[:-1] + (0.0,)))
    assert e.value.code == abi.E_INVALID
    for w in (wo, wg):  # a gear over two joints of the world: the defs and the created records agree; bad couples are refused
        w._r = w.create_joint(w.revolute_joint_def(w.body(0), w.body(1), (1.0, 1.0)))
        w._p = w.create_joint(w.prismatic_joint_def(w.body(1), w.body(0), (1.3, 1.2), (0.6, -0.8)))
        w._d = w.create_joint(w.distance_joint_def(w.body(0), w.body(1), (0.0, 1.0), (2.0, 1.0)))
    assert bytes(wo.gear_joint_def(wo._r, wo._p, -1.5)) == bytes(wg.gear_joint_def(wg._r, wg._p, -1.5))
    wo.create_joint(wo.gear_joint_def(wo._r, wo._p, -1.5))
    wg.create_joint(wg.gear_joint_def(wg._r, wg._p, -1.5))
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    with pytest.raises(B2gpuError) as e:  # a distance joint cannot be geared: the reference asserts
        wg.create_joint(wg.gear_joint_def(wg._r, wg._d, 1.0))
    assert e.value.code == abi.E_INVALID
    pg.lower_angle, pg.upper_angle = 1.0, 0.5  # lower > upper: the reference asserts
    with pytest.raises(B2gpuError) as e:
        wg.create_joint(pg)
    assert e.value.code == abi.E_INVALID
    jg.type = 11  # past the end of B2jointType
    with pytest.raises(B2gpuError) as e:
        wg.create_joint(jg)
    assert e.value.code == abi.E_UNSUPPORTED
    wg.close()
    ctx.close()


def test_oracle_joint_prevents_collision_unless_collide_connected(built):
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef
    from oracle import b2o
    counts = []
    for collide in (0, 1):
        w = b2o.B2world((0.0, 0.0))
        a = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 0.0)))
        a.create_fixture_by_shape(w.shapes.polygon_box(1.0, 1.0), 1.0)
        b = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(1.5, 0.0)))
        b.create_fixture_by_shape(w.shapes.polygon_box(1.0, 1.0), 1.0)
        jd = w.revolute_joint_def(a, b, (0.75, 0.0))
        jd.collide_connected = collide
        w.create_joint(jd)
        w.step(scenes.DT, 8, 3)
        counts.append(w.get_contact_count())
    assert counts == [0, 1]


# ---------------------------------------------------------------------------------------------- device stages vs oracle
def _free_running(name, ctx, every):
    from box2d_rs_b200 import scenes
    wo, wg, steps, ro, rg = _pair(name, ctx)
    assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
    assert wo.get_joint_count() == wg.get_joint_count() > 0
    for i in range(steps):
        if name == "joints_mix":  # user edits of the motorised arm mid-run (B2revoluteJoint setters)
            if i == 150:
                for m in (ro, rg):
                    m.set_motor_speed(-2.0)
            if i == 220:
                for m in (ro, rg):
                    m.enable_limit(False)
                    m.set_max_motor_torque(15.0)
            if i == 260:
                for m in (ro, rg):
                    m.enable_motor(False)
                    m.enable_limit(True)
                    m.set_limits(-0.5, 0.5)
        if name == "pulleys" and i in (100, 200):  # the drag moves on (B2mouseJoint::set_target)
            for m in (ro, rg):
                m.set_target((-26.0, 5.0) if i == 100 else (-34.0, 12.0))
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 3 or i % every == every - 1 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
            assert bad == [], "%s step %d: %s" % (name, i, bad[:6])
    assert np.abs(wo.snapshot().joints["impulse"]).max() > 0.0
    wg.close()


def _batched(name, ctx, n, lane_block):
    """Replicas of a joint scene in one batch, some perturbed: every memory-block size gives the oracle's worlds; flags
    (warm starting off, block solver off) and a sleeping world included."""
    from box2d_rs_b200 import scenes
    wo, wg, steps, _, _ = _pair(name, ctx)
    bt = wg.batch(n, lane_block=lane_block)
    picks = {0: wo.clone(), n - 1: wo.clone()}
    picks[n - 1].body(wo.get_body_count() - 1).set_linear_velocity((0.7, -0.3))
    picks[n - 1].set_warm_starting(False)
    for w, o in picks.items():
        bt.upload_world(w, o.snapshot())
    for i in range(min(steps, 160)):
        bt.step(scenes.DT, 8, 3)
        for o in picks.values():
            o.step(scenes.DT, 8, 3)
        if i % 40 == 39:
            for w, o in picks.items():
                bad = parity.compare_snapshots(o.snapshot(), bt.download_world(w)) + parity.compare_stats(o.get_stats(), bt.stats()[w])
                assert bad == [], "%s world %d step %d: %s" % (name, w, i, bad[:6])
    bt.close()
    wg.close()


def _batch_joint_controls(ctx, lane_block):
    """b2gpu_batch_set_joint_control: per-world motor speeds / torques / mouse targets of a batch (the RL action on a jointed
    agent) against oracle worlds driven through the single-world setters, sleeping worlds woken by a changed value included."""
    from box2d_rs_b200 import scenes
    from box2d_rs_b200.lib import B2gpuError
    n = 37
    rng = np.random.default_rng(5)
    for name, control in (("car", "speed"), ("joints_mix", "torque"), ("pulleys", "target")):
        wo, wg, steps, ro, rg = _pair(name, ctx)
        jo = ro[0] if name == "car" else ro     # the rear wheel's motor / the motorised arm / the mouse joint
        ji = jo.index
        bt = wg.batch(n, lane_block=lane_block)
        picks = {0: wo.clone(), 5: wo.clone(), n - 1: wo.clone()}
        for i in range(150):
            if i % 30 == 10:  # a new action for every world; world 5 repeats its old one every other time (no wake, no change)
                if control == "target":
                    vals = rng.uniform(-30.0, -20.0, (n, 2)).astype(np.float32) + np.float32([0.0, 30.0])
                else:
                    vals = rng.uniform(-30.0, 30.0, n).astype(np.float32) if control == "speed" else rng.uniform(0.0, 80.0, n).astype(np.float32)
                if i % 60 == 40:
                    vals[5] = last[5]
                last = vals.copy()
                if control == "speed":
                    bt.set_motor_speeds(ji, vals)
                elif control == "torque":
                    bt.set_max_motor_torques(ji, vals[:20])
                    bt.set_max_motor_torques(ji, vals[20:], first=20)
                else:
                    bt.set_targets(ji, vals)
                for w, o in picks.items():
                    j = o.joint(ji)
                    if control == "speed":
                        j.set_motor_speed(float(vals[w]))
                    elif control == "torque":
                        j.set_max_motor_torque(float(vals[w]))
                    else:
                        j.set_target((float(vals[w][0]), float(vals[w][1])))
            bt.step(scenes.DT, 8, 3)
            for o in picks.values():
                o.step(scenes.DT, 8, 3)
            if i % 10 == 9:
                for w, o in picks.items():
                    bad = parity.compare_snapshots(o.snapshot(), bt.download_world(w)) + parity.compare_stats(o.get_stats(), bt.stats()[w])
                    assert bad == [], "%s world %d step %d: %s" % (name, w, i, bad[:6])
        with pytest.raises(B2gpuError):  # a control the joint type does not have
            if control == "target":
                bt.set_motor_speeds(ji, np.zeros(n, np.float32))
            else:
                bt.set_targets(ji, np.zeros((n, 2), np.float32))
        bt.close()
        wg.close()


@pytest.mark.parametrize("lane_block", [1, 32])
def test_hostsim_batch_joint_controls(lane_block, hctx):
    _batch_joint_controls(hctx, lane_block)


@pytest.mark.gpu
def test_gpu_batch_joint_controls(gctx):
    _batch_joint_controls(gctx, 32)


def _teacher_forced(name, ctx):
    from box2d_rs_b200 import scenes
    wo, wg, steps, _, _ = _pair(name, ctx)
    bt = wg.batch(2, lane_block=1)
    for i in range(min(steps, 150)):
        if i % 5 == 0:
            bt.upload_world(1, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            bad = parity.compare_snapshots(wo.snapshot(), bt.download_world(1)) + parity.compare_stats(wo.get_stats(), bt.stats()[1])
            assert bad == [], "%s step %d: %s" % (name, i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    bt.close()
    wg.close()


@pytest.mark.parametrize("name", list(JOINT_SCENES))
def test_hostsim_free_running(name, hctx):
    _free_running(name, hctx, 20)


@pytest.mark.parametrize("name,lane_block", [("bridge", 1), ("tumbler", 4), ("joints_mix", 32), ("gears", 4), ("pulleys", 32)])
def test_hostsim_batched(name, lane_block, hctx):
    _batched(name, hctx, 5 if lane_block < 32 else 34, lane_block)


@pytest.mark.parametrize("name", list(JOINT_SCENES))
def test_hostsim_teacher_forced(name, hctx):
    _teacher_forced(name, hctx)


def _destroy_joints(ctx):
    """B2world::destroy_joint between steps: a mouse grab released mid-flight (the testbed's mouse_up), a bridge cut in the
    middle (later joints move down one index, handles follow), a hinge with collide_connected == false removed from two
    overlapping boxes, which start to collide."""
    from box2d_rs_b200 import abi, scenes
    from box2d_rs_b200.abi import BodyDef, FixtureDef
    wo, wg, _, _, _ = _pair("bridge", ctx)
    extra = []
    for w in (wo, wg):
        ground = w.body(0)
        crate = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(0.0, 8.0)))
        crate.create_fixture(FixtureDef(density=1.0, friction=0.4), w.shapes.polygon_box(0.75, 0.75))
        jd = w.mouse_joint_def(ground, crate, (0.25, 8.25))
        jd.length = 1000.0 * 2.25
        jd.stiffness, jd.damping = w.linear_stiffness(5.0, 0.7, ground, crate)
        mouse = w.create_joint(jd)
        mouse.set_target((6.0, 12.0))
        a = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-14.0, 14.0)))
        a.create_fixture(FixtureDef(density=1.0, friction=0.3), w.shapes.polygon_box(1.0, 0.25))
        b = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(-13.0, 14.0)))
        b.create_fixture(FixtureDef(density=1.0, friction=0.3), w.shapes.polygon_box(1.0, 0.25))
        hinge = w.create_joint(w.revolute_joint_def(a, b, (-13.5, 14.0)))  # overlapping boxes, no contact while hinged
        extra.append((mouse, hinge))
    n0 = wo.get_joint_count()
    for i in range(200):
        if i == 40:
            for w, (mouse, hinge) in zip((wo, wg), extra):
                w.destroy_joint(w.joint(10))     # cut the bridge
                assert mouse.index == n0 - 3 and hinge.index == n0 - 2
        if i == 70:
            for w, (mouse, hinge) in zip((wo, wg), extra):
                mouse.set_target((-6.0, 14.0))
        if i == 90:
            for w, (mouse, hinge) in zip((wo, wg), extra):
                w.destroy_joint(mouse)           # mouse up
                assert mouse.index == -1 and hinge.index == n0 - 3
        if i == 120:
            for w, (mouse, hinge) in zip((wo, wg), extra):
                w.destroy_joint(hinge)
            assert wo.get_joint_count() == wg.get_joint_count() == n0 - 3
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i in (119, 199):
            so = wo.snapshot()
            nb = len(so.bodies)
            between = [c for c in so.contacts if {int(so.fixtures[c["fixture_a"]]["body"]), int(so.fixtures[c["fixture_b"]]["body"])} == {nb - 2, nb - 1}]
            assert len(between) == (0 if i == 119 else 1)  # should_collide: no contact while hinged, one once the pair is found again
        if i % 5 == 4 or i in (40, 41, 90, 91, 120, 121):
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
            assert bad == [], "step %d: %s" % (i, bad[:6])
    wg.close()


def test_hostsim_destroy_joint(hctx):
    _destroy_joints(hctx)


@pytest.mark.gpu
def test_gpu_destroy_joint(gctx):
    _destroy_joints(gctx)


def test_sleeping_island_with_joints(hctx):
    """Two boxes resting apart on the ground, tied by a distance joint, fall asleep as ONE island (the joint propagates
    the island DFS); an impulse on one wakes both in the same step."""
    from box2d_rs_b200 import abi, scenes, world
    from box2d_rs_b200.abi import BodyDef, FixtureDef
    from oracle import b2o

    def build(w):
        ground = w.create_body(BodyDef())
        ground.create_fixture_by_shape(w.shapes.edge_two_sided((-20.0, 0.0), (20.0, 0.0)), 0.0)
        boxes = []
        for px in (-3.0, 3.0):
            b = w.create_body(BodyDef(type=abi.DYNAMIC_BODY, position=(px, 0.51)))
            b.create_fixture(FixtureDef(density=1.0, friction=0.5), w.shapes.polygon_box(0.5, 0.5))
            boxes.append(b)
        w.create_joint(w.distance_joint_def(boxes[0], boxes[1], (-3.0, 0.51), (3.0, 0.51)))
    wo = b2o.B2world((0.0, -10.0))
    build(wo)
    wg = world.B2world((0.0, -10.0), ctx=hctx)
    build(wg)
    slept_at = None
    for i in range(200):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + parity.compare_stats(wo.get_stats(), wg.get_stats())
        assert bad == [], "step %d: %s" % (i, bad[:6])
        if slept_at is None and int(wo.get_stats()["awake_bodies"]) == 0:
            slept_at = i
            assert int(wo.get_stats()["islands"]) == 1  # one island although the boxes do not touch
            for w in (wo, wg):
                w.body(1).apply_linear_impulse_to_center((0.0, 3.0), True)
        elif slept_at is not None and i == slept_at + 1:
            assert int(wo.get_stats()["awake_bodies"]) == 2  # the island DFS through the joint woke the other box
    assert slept_at is not None
    wg.close()


def test_checkpoint_round_trip_with_joints(hctx, tmp_path):
    from box2d_rs_b200 import checkpoint, scenes, world
    wo, wg, _, _, _ = _pair("joints_mix", hctx)
    for _ in range(60):
        wo.step(scenes.DT, 8, 3)
    snap = wo.snapshot()
    path = str(tmp_path / "joints.b2snap")
    checkpoint.save(snap, path, hctx.L)
    back = checkpoint.load(path, hctx.L)
    assert parity.compare_snapshots(snap, back) == []
    w2 = world.B2world((0.0, -10.0), ctx=hctx)
    w2.upload(back)
    for _ in range(40):
        wo.step(scenes.DT, 8, 3)
        w2.step(scenes.DT, 8, 3)
    assert parity.compare_snapshots(wo.snapshot(), w2.snapshot()) == []
    w2.close()
    wg.close()


def test_validate_rejects_bad_joint_records(built):
    from box2d_rs_b200 import abi, checkpoint, lib, scenes
    from oracle import b2o

    def broken(mutate):
        w = b2o.B2world((0.0, -10.0))
        scenes.bridge(w)
        s = w.snapshot()
        mutate(s)
        with pytest.raises(lib.B2gpuError) as e:
            checkpoint.validate(s)
        assert e.value.code == abi.E_INVALID

    def body_out_of_range(s):
        s.joints[3]["body_b"] = len(s.bodies)

    def same_body(s):
        s.joints[3]["body_b"] = s.joints[3]["body_a"]

    def unknown_type(s):
        s.joints[0]["type"] = 11
    broken(body_out_of_range)
    broken(same_body)
    broken(unknown_type)


def _large_teacher_forced(name, ctx, every):
    """Large-world mode 1 with joints: every step is the oracle's step of the same state (joint edges in the union-find
    islands and in the per-island traversal, joint rows in the per-island sweeps)."""
    from box2d_rs_b200 import scenes
    wo, wg, steps, _, _ = _pair(name, ctx)
    bt = wg.batch(1, lane_block=1, solver='large')
    for i in range(min(steps, 200)):
        if i % every == 0:
            bt.upload_world(0, wo.snapshot())
            wo.step(scenes.DT, 8, 3)
            bt.step(scenes.DT, 8, 3)
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(0), wo.get_stats(), bt.stats()[0])
            assert bad == [], "%s step %d: %s" % (name, i, bad[:6])
        else:
            wo.step(scenes.DT, 8, 3)
    bt.close()
    wg.close()


def _large_exact_free_running(name, ctx):
    """Large-world mode 2 (reference contact order) with joints: free-running state bit-identical to the oracle."""
    from box2d_rs_b200 import scenes
    wo, wg, steps, _, _ = _pair(name, ctx)
    wg.set_large_mode(2)
    for i in range(steps):
        wo.step(scenes.DT, 8, 3)
        wg.step(scenes.DT, 8, 3)
        if i < 2 or i % 40 == 39 or i == steps - 1:
            bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot()) + \
                [b for b in parity.compare_stats(wo.get_stats(), wg.get_stats()) if "island_bodies" not in b]
            assert bad == [], "%s step %d: %s" % (name, i, bad[:6])
    wg.close()


@pytest.mark.parametrize("name,every", [("bridge", 3), ("tumbler", 4), ("joints_mix", 2), ("gears", 2), ("pulleys", 3), ("car", 5), ("top_down", 2)])
def test_hostsim_large_mode_teacher_forced(name, every, hctx):
    _large_teacher_forced(name, hctx, every)


@pytest.mark.parametrize("name", list(JOINT_SCENES))
def test_hostsim_large_exact_free_running(name, hctx):
    _large_exact_free_running(name, hctx)


# ---------------------------------------------------------------------------------------------- CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(JOINT_SCENES))
def test_gpu_free_running(name, gctx):
    _free_running(name, gctx, 20)


@pytest.mark.gpu
@pytest.mark.parametrize("name,lane_block", [("bridge", 1), ("tumbler", 32), ("joints_mix", 32), ("gears", 32), ("pulleys", 32)])
def test_gpu_batched(name, lane_block, gctx):
    _batched(name, gctx, 5 if lane_block < 32 else 70, lane_block)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(JOINT_SCENES))
def test_gpu_teacher_forced(name, gctx):
    _teacher_forced(name, gctx)


@pytest.mark.gpu
@pytest.mark.parametrize("name,every", [("bridge", 3), ("tumbler", 4), ("joints_mix", 2)])
def test_gpu_large_mode_teacher_forced(name, every, gctx):
    _large_teacher_forced(name, gctx, every)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(JOINT_SCENES))
def test_gpu_large_exact_free_running(name, gctx):
    _large_exact_free_running(name, gctx)
