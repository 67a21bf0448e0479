#!/usr/bin/env python
"""Differential fuzzing of the device stages (host simulator build by default, --gpu for the CUDA path) against
the CPU oracle: random scenes (polygons of 3..8 vertices, circles, edges, chains, kinematic / fixed-rotation /
multi-fixture bodies, sensors, filters, restitution, damping, revolute / prismatic / wheel / distance / weld / friction / motor / pulley / mouse / gear joints with limits, motors and springs,
random world flags and iteration counts, dt = 0 steps,
mid-run set_transform / set_linear_velocity / apply_force / apply_torque / apply_*_impulse / set_awake edits), stepped freely and compared bit for bit.

  python tools/fuzz_parity.py --seeds 200 [--gpu] [--batch] [--large]

--large: the large-world mode (b2g_large.h), teacher-forced: every step starts from the oracle's state and must
reproduce the oracle's step (contacts created in that step compared as a set).
"""
import argparse
import os

EVERY = int(os.environ.get("FUZZ_EVERY", "16"))  # compare every n-th step (1 to locate a divergence)
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from box2d_rs_b200 import abi, scenes  # noqa: E402
from box2d_rs_b200.abi import BodyDef, FixtureDef  # noqa: E402


def build(world, rng):
    f32 = scenes.f32
    ground = world.create_body(BodyDef())
    body_types = [abi.STATIC_BODY]
    kind = rng.integers(0, 3)
    if kind == 0:
        ground.create_fixture_by_shape(world.shapes.edge_two_sided((-15.0, 0.0), (15.0, 0.0)), 0.0)
        ground.create_fixture_by_shape(world.shapes.edge_two_sided((-15.0, 0.0), (-15.0, 12.0)), 0.0)
        ground.create_fixture_by_shape(world.shapes.edge_two_sided((15.0, 0.0), (15.0, 12.0)), 0.0)
    elif kind == 1:
        pts = [(15.0, 10.0), (12.0, 1.0), (4.0, f32(rng.uniform(-1, 1))), (-4.0, f32(rng.uniform(-1, 1))), (-12.0, 1.0), (-15.0, 10.0)]
        ground.create_fixture(FixtureDef(friction=f32(rng.uniform(0, 1))), world.shapes.chain(pts, (16.0, 11.0), (-16.0, 11.0)))
    else:
        ground.create_fixture_by_shape(world.shapes.polygon_box(16.0, 0.5, (0.0, -0.5), f32(rng.uniform(-0.1, 0.1))), 0.0)
        ground.create_fixture_by_shape(world.shapes.circle(2.0, (f32(rng.uniform(-5, 5)), 0.0)), 0.0)
    n = int(rng.integers(2, 40))
    for k in range(n):
        t = rng.integers(0, 10)
        bd = BodyDef(type=abi.DYNAMIC_BODY, position=(f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))),
                     angle=f32(rng.uniform(-4, 4)), linear_velocity=(f32(rng.uniform(-3, 3)), f32(rng.uniform(-3, 3))),
                     angular_velocity=f32(rng.uniform(-2, 2)))
        if t == 0:
            bd.type = abi.KINEMATIC_BODY
        if t == 1:
            bd.fixed_rotation = 1
        if t == 2:
            bd.linear_damping, bd.angular_damping = f32(rng.uniform(0, 1)), f32(rng.uniform(0, 1))
        if t == 3:
            bd.gravity_scale = f32(rng.uniform(-0.5, 2))
        if t == 4:
            bd.allow_sleep = 0
        if t == 5:
            bd.awake = 0
        if t == 6 and rng.integers(0, 3) == 0:
            bd.enabled = 0  # no proxies: the fixture never collides
        if t == 7:  # fast spinner / fast mover: rotation and translation clamps, sin/cos range reduction
            bd.angular_velocity = f32(rng.uniform(-120, 120))
            bd.linear_velocity = (f32(rng.uniform(-150, 150)), f32(rng.uniform(-50, 150)))
            bd.angle = f32(rng.uniform(-200, 200))
        b = world.create_body(bd)
        body_types.append(bd.type if bd.enabled else -1)  # -1: disabled (its joints stay out of the islands)
        for _ in range(1 + int(rng.integers(0, 3) == 0)):
            zero_mass = bd.type == abi.DYNAMIC_BODY and rng.integers(0, 25) == 0  # dynamic body nothing can push
            fd = FixtureDef(density=f32(rng.uniform(0.2, 4)) if (bd.type == abi.DYNAMIC_BODY and not zero_mass) else 0.0,
                            friction=f32(rng.uniform(0, 1)), restitution=f32(rng.uniform(0, 0.8)) if rng.integers(0, 3) == 0 else 0.0)
            r = rng.integers(0, 8)
            if r == 0:
                fd.group_index = int(rng.integers(-2, 3))
            if r == 1:
                fd.category_bits, fd.mask_bits = int(1 << rng.integers(0, 3)), int(0xFFFF ^ (1 << rng.integers(0, 3)))
            if r == 2:
                fd.is_sensor = 1  # touching by GJK overlap, no manifold, no response
            s = rng.integers(0, 4)
            if s == 0:
                shape = world.shapes.circle(f32(rng.uniform(0.1, 0.6)), (f32(rng.uniform(-0.3, 0.3)), f32(rng.uniform(-0.3, 0.3))))
            elif s == 1:
                shape = world.shapes.polygon_box(f32(rng.uniform(0.1, 0.8)), f32(rng.uniform(0.1, 0.8)))
            elif s == 2:
                shape = world.shapes.polygon_box(f32(rng.uniform(0.1, 0.6)), f32(rng.uniform(0.1, 0.6)),
                                                 (f32(rng.uniform(-0.3, 0.3)), f32(rng.uniform(-0.3, 0.3))), f32(rng.uniform(-3, 3)))
            else:
                nv = int(rng.integers(3, 9))
                rad = rng.uniform(0.2, 0.7)
                ang = np.sort(rng.uniform(0, 2 * math.pi, nv))
                shape = world.shapes.polygon([(f32(rad * math.cos(a)), f32(rad * math.sin(a) * rng.uniform(0.5, 1.0))) for a in ang])
            b.create_fixture(fd, shape)
    # joints (all ten types) between random bodies, the ground included: limits, motors, soft springs, slack ranges,
    # collide_connected, degenerate pairs (kinematic or fixed-rotation bodies, anchors far from the bodies)
    joints, ends, gears, coupled = [], [], [], []
    if rng.integers(0, 5) < 3:
        for _ in range(int(rng.integers(1, 9))):
            a, b = int(rng.integers(0, n + 1)), int(rng.integers(0, n + 1))
            if a == b:
                continue
            kind = int(rng.integers(0, 9))
            if kind == 8:  # mouse joint: soft drag to a world target (sometimes rigid: stiffness 0 -> gamma 0), weak or strong
                jd = world.mouse_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))))
                jd.length = f32(rng.uniform(0, 400))
                if rng.integers(0, 4):
                    jd.stiffness, jd.damping = f32(rng.uniform(0, 200)), f32(rng.uniform(0, 20))
            elif kind == 7:  # pulley: random anchors and ground anchors (sometimes on top of each other), ratios around 1
                pa, pb = (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 10))), (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 10)))
                ga = pa if rng.integers(0, 6) == 0 else (f32(pa[0] + rng.uniform(-2, 2)), f32(pa[1] + rng.uniform(0, 8)))
                gb = (f32(pb[0] + rng.uniform(-2, 2)), f32(pb[1] + rng.uniform(0, 8)))
                jd = world.pulley_joint_def(a, b, ga, gb, pa, pb, f32(rng.uniform(0.3, 3.0)))
                jd.collide_connected = int(rng.integers(0, 2))
            elif kind == 6:  # motor joint: target pose, force / torque caps, correction factor
                jd = world.motor_joint_def(a, b)
                jd.local_anchor_a[0] = f32(jd.local_anchor_a[0] + rng.uniform(-2, 2))
                jd.local_anchor_a[1] = f32(jd.local_anchor_a[1] + rng.uniform(-2, 2))
                jd.reference_angle = f32(jd.reference_angle + rng.uniform(-1, 1))
                jd.length, jd.max_motor_torque = f32(rng.uniform(0, 300)), f32(rng.uniform(0, 300))
                jd.stiffness = f32(rng.uniform(0.0, 1.0))
            elif kind == 5:  # friction joint
                jd = world.friction_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))))
                jd.length, jd.max_motor_torque = f32(rng.uniform(0, 60)), f32(rng.uniform(0, 30))
            elif kind == 4:  # wheel: random (unnormalised) axis, spring, limits, motor
                th = rng.uniform(0, 2 * math.pi)
                jd = world.wheel_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))),
                                           (f32(math.cos(th) * rng.uniform(0.5, 2)), f32(math.sin(th) * rng.uniform(0.5, 2))))
                if rng.integers(0, 3) > 0:
                    jd.stiffness, jd.damping = world.linear_stiffness(f32(rng.uniform(0.5, 6)), f32(rng.uniform(0, 1)), a, b)
                if rng.integers(0, 2) == 0:
                    lo = f32(rng.uniform(-2, 0.2))
                    jd.enable_limit, jd.lower_angle, jd.upper_angle = 1, lo, f32(lo + (0.0 if rng.integers(0, 5) == 0 else rng.uniform(0, 3)))
                if rng.integers(0, 2) == 0:
                    jd.enable_motor, jd.motor_speed, jd.max_motor_torque = 1, f32(rng.uniform(-10, 10)), f32(rng.uniform(0, 100))
            elif kind == 3:  # prismatic: random axis, limits (sometimes equal), motor
                th = rng.uniform(0, 2 * math.pi)
                jd = world.prismatic_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))),
                                               (f32(math.cos(th) * rng.uniform(0.5, 2)), f32(math.sin(th) * rng.uniform(0.5, 2))))
                if rng.integers(0, 2) == 0:
                    lo = f32(rng.uniform(-3, 0.5))
                    jd.enable_limit, jd.lower_angle, jd.upper_angle = 1, lo, f32(lo + (0.0 if rng.integers(0, 5) == 0 else rng.uniform(0, 4)))
                if rng.integers(0, 2) == 0:
                    jd.enable_motor, jd.motor_speed, jd.max_motor_torque = 1, f32(rng.uniform(-3, 3)), f32(rng.uniform(0, 500))
            elif kind == 2:  # weld: rigid or soft; anchors far from the bodies and immovable partners included
                jd = world.weld_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))))
                if rng.integers(0, 2) == 0:
                    jd.stiffness, jd.damping = world.angular_stiffness(f32(rng.uniform(0.5, 8)), f32(rng.uniform(0, 1)), a, b)
            elif kind == 0:
                jd = world.revolute_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))))
                if rng.integers(0, 2) == 0:
                    lo = f32(rng.uniform(-1.5, 0.2))
                    jd.enable_limit, jd.lower_angle, jd.upper_angle = 1, lo, f32(lo + (0.0 if rng.integers(0, 5) == 0 else rng.uniform(0, 2)))
                if rng.integers(0, 2) == 0:
                    jd.enable_motor, jd.motor_speed, jd.max_motor_torque = 1, f32(rng.uniform(-3, 3)), f32(rng.uniform(0, 200))
            else:
                jd = world.distance_joint_def(a, b, (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))),
                                              (f32(rng.uniform(-10, 10)), f32(rng.uniform(0.5, 14))))
                r = rng.integers(0, 4)
                if r == 0:
                    jd.stiffness, jd.damping = world.linear_stiffness(f32(rng.uniform(0.5, 5)), f32(rng.uniform(0, 1)), a, b)
                if r <= 1:
                    jd.min_length = f32(max(jd.length - rng.uniform(0, 2), 0.0))
                    jd.max_length = f32(jd.length + rng.uniform(0, 2))
            jd.collide_connected = int(rng.integers(0, 2))
            joints.append((world.create_joint(jd), jd.type))
            ends.append((a, b))
        # gear joints over the revolute / prismatic joints made above (body B of each must be dynamic): any ratio sign, the
        # joint's own bodies or — as the testbed does — other ones, a joint geared to itself's neighbour on a shared body
        # (all four bodies must reach the gear's island through the coupled joints — a disabled one would leave the reference
        # with a stale island index — so the optional other bodies are the coupled joints' own body A)
        couples = [q for q, (h, t) in enumerate(joints) if t in (abi.JOINT_REVOLUTE, abi.JOINT_PRISMATIC)
                   and body_types[ends[q][1]] == abi.DYNAMIC_BODY and body_types[ends[q][0]] >= 0]
        for _ in range(int(rng.integers(0, 3))):
            if len(couples) < 2:
                break
            q1, q2 = (int(v) for v in rng.choice(couples, 2, replace=False))
            jd = world.gear_joint_def(joints[q1][0], joints[q2][0], f32(rng.uniform(0.3, 3.0) * (1 if rng.integers(0, 2) else -1)))
            if rng.integers(0, 4) == 0:
                jd.body_a = ends[q1][0]
            if rng.integers(0, 4) == 0:
                jd.body_b = ends[q2][0]
            if jd.body_a == jd.body_b:
                continue
            jd.collide_connected = int(rng.integers(0, 2))
            gears.append((world.create_joint(jd), jd.type))
            coupled += [joints[q1][0], joints[q2][0]]
    world._fuzz_joints = joints + gears
    world._fuzz_coupled = coupled  # never destroyed: "destroy the gear joint first" (src/joints/b2_gear_joint.rs:116-117)
    return n + 1


def run_seed(seed, make_world, steps, batch_mode, large=False, events=False, level_threshold=0):
    import parity
    from oracle import b2o
    rng = np.random.default_rng(seed)
    g = (0.0, f32v(rng.uniform(-12, 0)))
    wo = b2o.B2world(g)
    wg = make_world(g)
    state = rng.bit_generator.state
    try:
        nb = build(wo, rng)
    except ValueError:
        wg.close()
        return "skip"
    rng.bit_generator.state = state
    build(wg, rng)
    flags = (bool(rng.integers(0, 4)), bool(rng.integers(0, 4)), bool(rng.integers(0, 4)))
    for w in (wo, wg):
        w.set_warm_starting(flags[0]); w.set_block_solve(flags[1]); w.set_allow_sleeping(flags[2])
    vi, pi = int(rng.integers(1, 10)), int(rng.integers(0, 5))
    bad = parity.compare_snapshots(wo.snapshot(), wg.snapshot())
    if bad:
        return "setup: %s" % bad[:3]
    stepper = wg
    batch = None
    if batch_mode:
        batch = wg.batch(int(rng.integers(33, 70)), max_contacts=40 * nb + 256)  # the reference grows its tables; a batch cannot
    if large == 2:
        # mode 2 (replica tree kept): FREE-RUNNING, every table including the tree compared; the state is re-uploaded only after
        # a user edit of the oracle world (a batch has no set_transform)
        bt = wg.batch(1, lane_block=1, solver='large_exact')
        bt.upload_world(0, wo.snapshot())
        for i in range(steps):
            dt = 0.0 if rng.integers(0, 40) == 0 else scenes.DT
            ev = rng.integers(0, 30)
            if ev <= 1:
                b = wo.body(int(rng.integers(1, nb)))
                if ev == 0:
                    b.set_transform((f32v(rng.uniform(-8, 8)), f32v(rng.uniform(1, 10))), f32v(rng.uniform(-3, 3)))
                else:
                    b.set_linear_velocity((f32v(rng.uniform(-6, 6)), f32v(rng.uniform(-6, 6))))
                bt.upload_world(0, wo.snapshot())
            wo.step(dt, vi, pi)
            if exploded(wo):
                run_seed.exploded = getattr(run_seed, "exploded", 0) + 1
                break
            bt.step(dt, vi, pi)
            if i % EVERY == EVERY - 1 or i == steps - 1:
                bad = parity.compare_snapshots(wo.snapshot(), bt.download_world(0))
                bad += [b for b in parity.compare_stats(wo.get_stats(), bt.stats()[0]) if "island_bodies" not in b]
                if bad:
                    return "large exact, step %d (dt=%g vi=%d pi=%d flags=%s): %s" % (i, dt, vi, pi, flags, bad[:4])
        bt.close()
        wg.close()
        return None
    if large:
        bt = wg.batch(1, lane_block=1, solver='large')
        if level_threshold:
            bt.set_level_threshold(level_threshold)
        for i in range(steps):
            dt = 0.0 if rng.integers(0, 40) == 0 else scenes.DT
            ev = rng.integers(0, 30)
            if ev == 0:
                wo.body(int(rng.integers(1, nb))).set_transform((f32v(rng.uniform(-8, 8)), f32v(rng.uniform(1, 10))), f32v(rng.uniform(-3, 3)))
            if ev == 1:
                wo.body(int(rng.integers(1, nb))).set_linear_velocity((f32v(rng.uniform(-6, 6)), f32v(rng.uniform(-6, 6))))
            bt.upload_world(0, wo.snapshot())
            wo.step(dt, vi, pi)
            if exploded(wo):
                run_seed.exploded = getattr(run_seed, "exploded", 0) + 1
                break
            bt.step(dt, vi, pi)
            bad = parity.compare_large_step(wo.snapshot(), bt.download_world(0), wo.get_stats(), bt.stats()[0])
            if bad:
                return "large, step %d (dt=%g vi=%d pi=%d flags=%s): %s" % (i, dt, vi, pi, flags, bad[:4])
        bt.close()
        wg.close()
        return None
    for i in range(steps):
        dt = 0.0 if rng.integers(0, 40) == 0 else scenes.DT
        ev = rng.integers(0, 30)
        if os.environ.get("FUZZ_TRACE"):
            print("step %d: dt %g event %d" % (i, dt, ev))
        if batch is None and ev == 0:
            b = int(rng.integers(1, nb))
            p, a = (f32v(rng.uniform(-8, 8)), f32v(rng.uniform(1, 10))), f32v(rng.uniform(-3, 3))
            wo.body(b).set_transform(p, a); wg.body(b).set_transform(p, a)
        if batch is None and ev == 1:
            b = int(rng.integers(1, nb))
            v = (f32v(rng.uniform(-6, 6)), f32v(rng.uniform(-6, 6)))
            wo.body(b).set_linear_velocity(v); wg.body(b).set_linear_velocity(v)
        if batch is None and ev == 8 and wo._fuzz_joints:  # B2revoluteJoint setters mid-run
            q = int(rng.integers(0, len(wo._fuzz_joints)))
            if wo._fuzz_joints[q][1] == abi.JOINT_MOUSE:  # B2mouseJoint::set_target
                tgt = (f32v(rng.uniform(-10, 10)), f32v(rng.uniform(0.5, 14)))
                for w in (wo, wg):
                    w._fuzz_joints[q][0].set_target(tgt)
            if wo._fuzz_joints[q][1] in (abi.JOINT_REVOLUTE, abi.JOINT_PRISMATIC, abi.JOINT_WHEEL):
                op, val, flag = int(rng.integers(0, 5)), f32v(rng.uniform(-3, 3)), bool(rng.integers(0, 2))
                for w in (wo, wg):
                    j = w._fuzz_joints[q][0]
                    if op == 0: j.set_motor_speed(val)
                    elif op == 1: j.set_max_motor_torque(abs(val) * 50.0)
                    elif op == 2: j.enable_motor(flag)
                    elif op == 3: j.enable_limit(flag)
                    else: j.set_limits(min(val, 0.0) - 0.3, max(val, 0.0) + 0.3)
        if batch is None and ev == 9 and wo._fuzz_joints and rng.integers(0, 3) == 0:  # B2world::destroy_joint
            q = int(rng.integers(0, len(wo._fuzz_joints)))
            if not any(wo._fuzz_joints[q][0] is h for h in wo._fuzz_coupled):
                for w in (wo, wg):
                    w.destroy_joint(w._fuzz_joints.pop(q)[0])
        if batch is None and 2 <= ev <= 7:  # the rest of B2body's force / impulse API, sleeping bodies included
            b = int(rng.integers(1, nb))
            vec = (f32v(rng.uniform(-40, 40)), f32v(rng.uniform(-40, 40)))
            pt = (f32v(rng.uniform(-8, 8)), f32v(rng.uniform(0, 10)))
            sc = f32v(rng.uniform(-20, 20))
            wake = bool(rng.integers(0, 3))
            for w in (wo, wg):
                bd = w.body(b)
                if ev == 2: bd.apply_force(vec, pt, wake)
                elif ev == 3: bd.apply_torque(sc, wake)
                elif ev == 4: bd.apply_linear_impulse((vec[0] * 0.1, vec[1] * 0.1), pt, wake)
                elif ev == 5: bd.apply_linear_impulse_to_center((vec[0] * 0.1, vec[1] * 0.1), wake)
                elif ev == 6: bd.apply_angular_impulse(sc * 0.05, wake)
                else: bd.set_awake(wake)
        wo.step(dt, vi, pi)
        if exploded(wo):
            run_seed.exploded = getattr(run_seed, "exploded", 0) + 1
            break
        if batch is None and events:  # begin / end contact events of every step, in the reference's firing order
            ev_g = wg.step_with_events(dt, vi, pi)
            ev_o = wo.contact_events()
            tab = np.stack([ev_g[k] for k in ("type", "fixture_a", "index_a", "fixture_b", "index_b")], 1) if len(ev_g) else np.zeros((0, 5), np.int32)
            if not np.array_equal(tab, ev_o):
                return "events, step %d: %s vs %s" % (i, tab[:4].tolist(), ev_o[:4].tolist())
            run_seed.event_total = getattr(run_seed, "event_total", 0) + len(ev_o)
        elif batch is None:
            wg.step(dt, vi, pi)
        else:
            batch.step(dt, vi, pi)
        if i % EVERY == EVERY - 1 or i == steps - 1:
            got = wg.snapshot() if batch is None else batch.download_world(batch.n_worlds - 1)
            st = wg.get_stats() if batch is None else batch.stats()[batch.n_worlds - 1]
            bad = parity.compare_snapshots(wo.snapshot(), got) + parity.compare_stats(wo.get_stats(), st)
            if bad:
                return "step %d (vi=%d pi=%d flags=%s): %s" % (i, vi, pi, flags, bad[:4])
    if batch is not None:
        batch.close()
    wg.close()
    return None


def f32v(x):
    return float(np.float32(x))


def exploded(world):
    """Random joints can over-constrain a scene until velocities overflow: once the oracle's state holds inf / NaN the run is
    over (0 * inf = NaN then reaches even static bodies in the reference, which the device never writes): not a parity case."""
    return not np.isfinite(world.body_state()).all()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--batch", action="store_true")
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--large-exact", action="store_true", help="large-world mode 2 (replica tree kept), free-running, the tree tables compared too")
    ap.add_argument("--level-threshold", type=int, default=0, help="--large: islands of at least this many contacts take the level-scheduled sweeps (b2g_levels.h)")
    ap.add_argument("--events", action="store_true", help="also compare b2gpu_contact_events with the oracle's listener log every step")
    args = ap.parse_args()
    from box2d_rs_b200 import batch as batch_mod, world
    lib_path = None if args.gpu else os.path.join(ROOT, "tests", "hostsim", "libb2gpu_hostsim.so")
    ctx = batch_mod.Context(0, lib_path=lib_path)
    fails = 0
    for seed in range(args.first, args.first + args.seeds):
        r = run_seed(seed, lambda g: world.B2world(g, ctx=ctx), args.steps, args.batch, 2 if args.large_exact else args.large, args.events, args.level_threshold)
        if r not in (None, "skip"):
            fails += 1
            print("seed %d: %s" % (seed, r), flush=True)
    print("fuzz: %d seeds, %d failures, %d runs ended early by a numerical explosion of the scene (%s%s%s)"
          % (args.seeds, fails, getattr(run_seed, "exploded", 0), "gpu" if args.gpu else "host simulator",
                                                     ", batch" if args.batch else ", large-world mode 2 free-running" if args.large_exact else (", large-world mode teacher-forced" + (", level threshold %d" % args.level_threshold if args.level_threshold else "")) if args.large else "",
                                                     ", %d contact events compared" % getattr(run_seed, "event_total", 0) if args.events else ""))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
