"""Mirror of the reference's world-building API (B2world / B2body / shapes) on top of the C ABI.

Method names and argument meaning follow box2d-rs (src/b2_world.rs, src/b2_body.rs, src/shapes/*):
B2world::new(gravity), create_body(&B2bodyDef), B2body::create_fixture(&B2fixtureDef),
create_fixture_by_shape(shape, density), set_transform, set_linear_velocity, B2world::step(dt,
velocity_iterations, position_iterations), set_allow_sleeping, set_warm_starting ...
Scene recipes written against this API (box2d_rs_b200/scenes.py) run unchanged on the GPU engine.
Stepping happens on the GPU only; on a machine without one, step() raises B2gpuError(NO_DEVICE).
"""
import ctypes as C

import numpy as np

from . import abi
from .batch import Batch, Context
from .lib import check


class Shapes:
    """Shape factory: B2polygonShape::set_as_box / set_as_box_angle / set, B2circleShape, B2edgeShape, B2chainShape."""

    def __init__(self, L):
        self.L = L

    def polygon_box(self, hx, hy, center=None, angle=0.0):
        s = abi.ShapeDef()
        if center is None:
            check(self.L, self.L.b2gpu_polygon_set_as_box(C.byref(s), hx, hy))
        else:
            check(self.L, self.L.b2gpu_polygon_set_as_box_angle(C.byref(s), hx, hy, center[0], center[1], angle))
        return s

    def polygon(self, vertices):
        s = abi.ShapeDef()
        flat = (C.c_float * (2 * len(vertices)))(*[c for v in vertices for c in v])
        check(self.L, self.L.b2gpu_polygon_set(C.byref(s), flat, len(vertices)))
        return s

    circle = staticmethod(abi.circle_shape)
    edge_two_sided = staticmethod(abi.edge_two_sided)
    edge_one_sided = staticmethod(abi.edge_one_sided)
    chain = staticmethod(abi.chain_shape)

    def compute_mass(self, shape, density):
        md = abi.MassData()
        check(self.L, self.L.b2gpu_shape_compute_mass(C.byref(shape), density, C.byref(md)))
        return md


class B2body:
    def __init__(self, world, index):
        self.world, self.index = world, index

    def create_fixture(self, fixture_def, shape):
        w = self.world
        return check(w.L, w.L.b2gpu_body_create_fixture(w.h, self.index, C.byref(fixture_def), C.byref(shape)))

    def create_fixture_by_shape(self, shape, density):
        return self.create_fixture(abi.FixtureDef(density=density), shape)

    def set_transform(self, position, angle):
        w = self.world
        check(w.L, w.L.b2gpu_body_set_transform(w.h, self.index, position[0], position[1], angle))

    def set_linear_velocity(self, v):
        w = self.world
        check(w.L, w.L.b2gpu_body_set_linear_velocity(w.h, self.index, v[0], v[1]))

    def set_angular_velocity(self, omega):
        w = self.world
        check(w.L, w.L.b2gpu_body_set_angular_velocity(w.h, self.index, omega))

    def apply_force_to_center(self, f, wake=True):
        w = self.world
        check(w.L, w.L.b2gpu_body_apply_force_to_center(w.h, self.index, f[0], f[1], int(wake)))

    def apply_force(self, f, point, wake=True):
        w = self.world
        check(w.L, w.L.b2gpu_body_apply_force(w.h, self.index, f[0], f[1], point[0], point[1], int(wake)))

    def apply_torque(self, torque, wake=True):
        w = self.world
        check(w.L, w.L.b2gpu_body_apply_torque(w.h, self.index, torque, int(wake)))

    def apply_linear_impulse(self, impulse, point, wake=True):
        w = self.world
        check(w.L, w.L.b2gpu_body_apply_linear_impulse(w.h, self.index, impulse[0], impulse[1], point[0], point[1], int(wake)))

    def apply_linear_impulse_to_center(self, impulse, wake=True):
        w = self.world
        check(w.L, w.L.b2gpu_body_apply_linear_impulse_to_center(w.h, self.index, impulse[0], impulse[1], int(wake)))

    def apply_angular_impulse(self, impulse, wake=True):
        w = self.world
        check(w.L, w.L.b2gpu_body_apply_angular_impulse(w.h, self.index, impulse, int(wake)))

    def set_awake(self, flag):
        w = self.world
        check(w.L, w.L.b2gpu_body_set_awake(w.h, self.index, int(flag)))

    def set_damping(self, linear_damping, angular_damping):
        """B2body::set_linear_damping + set_angular_damping."""
        w = self.world
        check(w.L, w.L.b2gpu_body_set_damping(w.h, self.index, linear_damping, angular_damping))

    def set_gravity_scale(self, scale):
        w = self.world
        check(w.L, w.L.b2gpu_body_set_gravity_scale(w.h, self.index, scale))

    def set_sleeping_allowed(self, flag):
        w = self.world
        check(w.L, w.L.b2gpu_body_set_sleeping_allowed(w.h, self.index, int(flag)))

    def _rec(self):
        w = self.world
        out = np.zeros(1, abi.BODY_DTYPE)
        check(w.L, w.L.b2gpu_world_get_body(w.h, self.index, out.ctypes.data))
        return out[0]

    def get_position(self):
        r = self._rec()
        return float(r["xf"][0]), float(r["xf"][1])

    def get_angle(self):
        return float(self._rec()["a"])

    def get_linear_velocity(self):
        r = self._rec()
        return float(r["v"][0]), float(r["v"][1])

    def is_awake(self):
        return bool(int(self._rec()["flags"]) & abi.BODY_AWAKE)


class B2joint:
    """B2revoluteJoint / B2distanceJoint handle (setters of src/joints/b2_revolute_joint.rs:172-242)."""

    def __init__(self, world, index):
        self.world, self.index = world, index

    def set_motor_speed(self, speed):
        check(self.world.L, self.world.L.b2gpu_joint_set_motor_speed(self.world.h, self.index, speed))

    def set_max_motor_torque(self, torque):
        check(self.world.L, self.world.L.b2gpu_joint_set_max_motor_torque(self.world.h, self.index, torque))

    def enable_motor(self, flag):
        check(self.world.L, self.world.L.b2gpu_joint_enable_motor(self.world.h, self.index, int(flag)))

    def enable_limit(self, flag):
        check(self.world.L, self.world.L.b2gpu_joint_enable_limit(self.world.h, self.index, int(flag)))

    def set_limits(self, lower, upper):
        check(self.world.L, self.world.L.b2gpu_joint_set_limits(self.world.h, self.index, lower, upper))

    def set_target(self, target):
        """B2mouseJoint::set_target (src/joints/b2_mouse_joint.rs:114-119)."""
        check(self.world.L, self.world.L.b2gpu_joint_set_target(self.world.h, self.index, target[0], target[1]))

    def record(self):
        out = np.zeros(1, abi.JOINT_DTYPE)
        check(self.world.L, self.world.L.b2gpu_world_get_joint(self.world.h, self.index, out.ctypes.data))
        return out[0]


def _body_index(b):
    return b.index if hasattr(b, "index") else int(b)


class B2world:
    def __init__(self, gravity, ctx=None, device=0, lib_path=None):
        self.ctx = ctx or Context(device, lib_path=lib_path)
        self.L = self.ctx.L
        self.shapes = Shapes(self.L)
        self.h = C.c_void_p()
        check(self.L, self.L.b2gpu_world_create(self.ctx.h, gravity[0], gravity[1], C.byref(self.h)))
        self._joint_handles = []

    def close(self):
        if self.h:
            self.L.b2gpu_world_destroy(self.h)
            self.h = C.c_void_p()

    def create_body(self, body_def):
        return B2body(self, check(self.L, self.L.b2gpu_world_create_body(self.h, C.byref(body_def))))

    def body(self, index):
        return B2body(self, index)

    def revolute_joint_def(self, body_a, body_b, anchor):
        """B2revoluteJointDef::default() + initialize(body_a, body_b, anchor): edit the fields, then create_joint."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_revolute_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b),
                                                      anchor[0], anchor[1]))
        return d

    def distance_joint_def(self, body_a, body_b, anchor_a, anchor_b):
        """B2distanceJointDef::default() + initialize(b1, b2, anchor1, anchor2)."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_distance_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b),
                                                      anchor_a[0], anchor_a[1], anchor_b[0], anchor_b[1]))
        return d

    def prismatic_joint_def(self, body_a, body_b, anchor, axis):
        """B2prismaticJointDef::default() + initialize(body_a, body_b, anchor, axis).  In the shared def struct lower_angle /
        upper_angle are the translation limits, max_motor_torque the maximum motor force, (length, min_length) the local axis."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_prismatic_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b),
                                                       anchor[0], anchor[1], axis[0], axis[1]))
        return d

    def friction_joint_def(self, body_a, body_b, anchor):
        """B2frictionJointDef::default() + initialize(body_a, body_b, anchor): `length` is max_force, `max_motor_torque` max_torque."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_friction_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b),
                                                      anchor[0], anchor[1]))
        return d

    def pulley_joint_def(self, body_a, body_b, ground_a, ground_b, anchor_a, anchor_b, ratio):
        """B2pulleyJointDef::default() + initialize(...) (src/joints/b2_pulley_joint.rs:10-78); overlay in b2gpu.h."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_pulley_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), ground_a[0],
                                                    ground_a[1], ground_b[0], ground_b[1], anchor_a[0], anchor_a[1],
                                                    anchor_b[0], anchor_b[1], ratio))
        return d

    def gear_joint_def(self, joint1, joint2, ratio):
        """B2gearJointDef::default() with joint1, joint2 (revolute / prismatic handles) and ratio
        (src/joints/b2_gear_joint.rs:12-40); overlay in b2gpu.h."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_gear_joint_def(self.h, C.byref(d), joint1.index, joint2.index, ratio))
        return d

    def mouse_joint_def(self, body_a, body_b, target):
        """B2mouseJointDef::default() with `target` (src/joints/b2_mouse_joint.rs:8-21): set length (= max_force),
        stiffness, damping."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_mouse_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), target[0],
                                                   target[1]))
        return d

    def motor_joint_def(self, body_a, body_b):
        """B2motorJointDef::default() + initialize(body_a, body_b): local_anchor_a is linear_offset, reference_angle angular_offset,
        `length` max_force, `max_motor_torque` max_torque, `stiffness` correction_factor."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_motor_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b)))
        return d

    def wheel_joint_def(self, body_a, body_b, anchor, axis):
        """B2wheelJointDef::default() + initialize(body_a, body_b, anchor, axis): lower_angle / upper_angle are the translation
        limits, (length, min_length) the local axis; stiffness / damping as for a distance joint."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_wheel_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b),
                                                   anchor[0], anchor[1], axis[0], axis[1]))
        return d

    def weld_joint_def(self, body_a, body_b, anchor):
        """B2weldJointDef::default() + initialize(body_a, body_b, anchor): rigid unless stiffness / damping are set."""
        d = abi.JointDef()
        check(self.L, self.L.b2gpu_weld_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b),
                                                  anchor[0], anchor[1]))
        return d

    def angular_stiffness(self, frequency_hertz, damping_ratio, body_a, body_b):
        """b2_angular_stiffness: (stiffness, damping) of a soft weld joint."""
        k, d = C.c_float(), C.c_float()
        check(self.L, self.L.b2gpu_angular_stiffness(self.h, frequency_hertz, damping_ratio, _body_index(body_a),
                                                     _body_index(body_b), C.byref(k), C.byref(d)))
        return k.value, d.value

    def linear_stiffness(self, frequency_hertz, damping_ratio, body_a, body_b):
        """b2_linear_stiffness: (stiffness, damping) of a soft distance joint."""
        k, d = C.c_float(), C.c_float()
        check(self.L, self.L.b2gpu_linear_stiffness(self.h, frequency_hertz, damping_ratio, _body_index(body_a),
                                                    _body_index(body_b), C.byref(k), C.byref(d)))
        return k.value, d.value

    def create_joint(self, joint_def):
        """B2world::create_joint (all ten joint types)."""
        j = B2joint(self, check(self.L, self.L.b2gpu_world_create_joint(self.h, C.byref(joint_def))))
        self._joint_handles.append(j)
        return j

    def destroy_joint(self, joint):
        """B2world::destroy_joint (src/private/dynamics/b2_world.rs:278-339): wakes both bodies; joints created later move
        down one index (handles made by create_joint follow)."""
        i = joint.index
        check(self.L, self.L.b2gpu_world_destroy_joint(self.h, i))
        self._joint_handles = [h for h in self._joint_handles if h is not joint]
        joint.index = -1
        for h in self._joint_handles:
            if h.index > i:
                h.index -= 1

    def joint(self, index):
        return B2joint(self, index)

    def get_joint_count(self):
        return check(self.L, self.L.b2gpu_world_get_joint_count(self.h))

    def set_allow_sleeping(self, flag):
        check(self.L, self.L.b2gpu_world_set_allow_sleeping(self.h, int(flag)))

    def set_gravity(self, gravity):
        """B2world::set_gravity."""
        check(self.L, self.L.b2gpu_world_set_gravity(self.h, gravity[0], gravity[1]))

    def get_gravity(self):
        gx, gy = C.c_float(), C.c_float()
        check(self.L, self.L.b2gpu_world_get_gravity(self.h, C.byref(gx), C.byref(gy)))
        return (gx.value, gy.value)

    def set_warm_starting(self, flag):
        check(self.L, self.L.b2gpu_world_set_warm_starting(self.h, int(flag)))

    def set_continuous_physics(self, flag):
        check(self.L, self.L.b2gpu_world_set_continuous_physics(self.h, int(flag)))

    def set_block_solve(self, flag):
        check(self.L, self.L.b2gpu_world_set_block_solve(self.h, int(flag)))

    def set_large_mode(self, flag):
        """Data-parallel ordered stages for one large world (b2gpu_world_set_large_mode): same step semantics,
        contacts created in one update_pairs call are appended in LBVH order instead of reference-tree order.
        flag = 2 keeps the replica tree (sequential re-insertion): reference contact order, bit-identical free-running."""
        check(self.L, self.L.b2gpu_world_set_large_mode(self.h, int(flag)))

    def set_level_threshold(self, contacts):
        """Large-world mode: islands with at least `contacts` contacts (and no joints) are swept by one CTA in dependency-level
        order (b2gpu_world_set_level_threshold); 0 = library default (1024), negative = never."""
        check(self.L, self.L.b2gpu_world_set_level_threshold(self.h, int(contacts)))

    def step(self, dt, velocity_iterations, position_iterations):
        check(self.L, self.L.b2gpu_world_step(self.h, dt, velocity_iterations, position_iterations))

    def get_body_count(self):
        return check(self.L, self.L.b2gpu_world_get_body_count(self.h))

    def get_contact_count(self):
        return check(self.L, self.L.b2gpu_world_get_contact_count(self.h))

    def get_stats(self):
        out = np.zeros(1, abi.STATS_DTYPE)
        check(self.L, self.L.b2gpu_world_get_stats(self.h, out.ctypes.data))
        return out[0]

    def snapshot(self):
        """Full step state (b2gpu_world_download) as abi.Snapshot."""
        n = abi.SnapshotSizes()
        check(self.L, self.L.b2gpu_world_snapshot_sizes(self.h, C.byref(n)))
        snap = abi.Snapshot(n)
        c = snap.as_c()
        check(self.L, self.L.b2gpu_world_download(self.h, C.byref(c)))
        return snap.finish(c)

    def upload(self, snap):
        c = snap.as_c()
        check(self.L, self.L.b2gpu_world_upload(self.h, C.byref(c)))

    def step_with_events(self, dt, velocity_iterations, position_iterations):
        """One step plus the begin_contact / end_contact events the reference's listener would have received during
        it, in firing order (b2gpu_contact_events; abi.CONTACT_EVENT_DTYPE).  Costs two snapshot downloads."""
        before = self.snapshot()
        self.step(dt, velocity_iterations, position_iterations)
        after = self.snapshot()
        return contact_events(self.L, before, after, int(self.get_stats()["destroyed"]))

    def post_solve_events(self):
        """B2contactListener::post_solve reports of the last step in the reference's call order (abi.POST_SOLVE_DTYPE)."""
        n = check(self.L, self.L.b2gpu_world_post_solve_events(self.h, None, 0))
        out = np.zeros(max(n, 1), abi.POST_SOLVE_DTYPE)
        check(self.L, self.L.b2gpu_world_post_solve_events(self.h, out.ctypes.data, n))
        return out[:n]

    def save_checkpoint(self, path):
        """Full step state to a snapshot file (b2gpu_snapshot_save); see checkpoint.py."""
        from . import checkpoint
        checkpoint.save(self.snapshot(), path, self.L)

    def load_checkpoint(self, path):
        """Resume from a snapshot file: the next step continues the saved run bit for bit."""
        from . import checkpoint
        self.upload(checkpoint.load(path, self.L))

    def ray_cast_closest(self, p1p2):
        """B2world::ray_cast with the closest-hit callback for an [n][4] array of rays (p1.x p1.y p2.x p2.y).
        Returns a structured array (abi.RAY_HIT_DTYPE); fixture == -1 where nothing was hit."""
        rays = np.ascontiguousarray(p1p2, np.float32).reshape(-1, 4)
        out = np.zeros(rays.shape[0], abi.RAY_HIT_DTYPE)
        check(self.L, self.L.b2gpu_world_ray_cast_closest(self.h, rays.ctypes.data, rays.shape[0], out.ctypes.data))
        return out

    def query_aabb(self, aabbs, max_hits=64):
        """B2world::query_aabb for an [n][4] array of boxes: list of [(fixture, child), ...] per box, report order."""
        boxes = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 4)
        counts = np.zeros(boxes.shape[0], np.int32)
        hits = np.zeros((boxes.shape[0], max(max_hits, 1), 2), np.int32)
        check(self.L, self.L.b2gpu_world_query_aabb(self.h, boxes.ctypes.data, boxes.shape[0], max_hits, counts.ctypes.data,
                                                    hits.ctypes.data))
        return [[(int(f), int(c)) for f, c in hits[i, :min(int(counts[i]), max_hits)]] for i in range(boxes.shape[0])], counts

    def batch(self, n_worlds, **kw):
        """n_worlds replicas of this world's current state, one CTA lane per world (b2gpu_batch_create)."""
        return Batch(self.ctx, self.snapshot(), n_worlds, **kw)


def contact_events(L, before, after, destroyed=-1):
    """b2gpu_contact_events on two abi.Snapshot objects -> structured array of abi.CONTACT_EVENT_DTYPE."""
    cb, ca = before.as_c(), after.as_c()
    n = check(L, L.b2gpu_contact_events(C.byref(cb), C.byref(ca), destroyed, None, 0))
    out = np.zeros(max(n, 1), abi.CONTACT_EVENT_DTYPE)
    check(L, L.b2gpu_contact_events(C.byref(cb), C.byref(ca), destroyed, out.ctypes.data, n))
    return out[:n]
