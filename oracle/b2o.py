"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU oracle (oracle/libb2o.so).

Only tests/, bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke()
may import this module.  PARITY UNPINNED beyond the reference's own four tests.
The class/method names mirror the reference API (B2world::create_body, B2body::create_fixture,
B2world::step, ...) so scene recipes run unchanged on the oracle and on the GPU engine.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from box2d_rs_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libb2o.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "b2gpu.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libb2o.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.b2o_world_create.restype = C.c_void_p
        L.b2o_world_create.argtypes = [C.c_float, C.c_float]
        L.b2o_world_clone.restype = C.c_void_p
        L.b2o_world_clone.argtypes = [C.c_void_p]
        L.b2o_world_destroy.argtypes = [C.c_void_p]
        L.b2o_create_body.argtypes = [C.c_void_p, C.POINTER(abi.BodyDef)]
        L.b2o_create_fixture.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.FixtureDef), C.POINTER(abi.ShapeDef)]
        L.b2o_set_transform.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.b2o_set_linear_velocity.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.b2o_set_angular_velocity.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.b2o_apply_force_to_center.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.b2o_apply_force.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.b2o_apply_linear_impulse.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.b2o_apply_linear_impulse_to_center.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.b2o_apply_torque.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.b2o_apply_angular_impulse.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.b2o_body_set_awake.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2o_body_set_damping.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.b2o_body_set_gravity_scale.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.b2o_body_set_sleeping_allowed.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("b2o_set_allow_sleeping", "b2o_set_warm_starting", "b2o_set_block_solve", "b2o_set_collect_levels"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.b2o_revolute_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float]
        L.b2o_distance_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float,
                                             C.c_float, C.c_float]
        L.b2o_prismatic_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float,
                                              C.c_float, C.c_float]
        L.b2o_friction_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float]
        L.b2o_motor_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int]
        L.b2o_pulley_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int] + [C.c_float] * 9
        L.b2o_gear_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float]
        L.b2o_mouse_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float]
        L.b2o_joint_set_target.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.b2o_wheel_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float,
                                          C.c_float, C.c_float]
        L.b2o_weld_joint_def.argtypes = [C.c_void_p, C.POINTER(abi.JointDef), C.c_int, C.c_int, C.c_float, C.c_float]
        L.b2o_angular_stiffness.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float)]
        L.b2o_linear_stiffness.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_float),
                                           C.POINTER(C.c_float)]
        L.b2o_create_joint.argtypes = [C.c_void_p, C.POINTER(abi.JointDef)]
        L.b2o_joint_count.argtypes = [C.c_void_p]
        L.b2o_destroy_joint.argtypes = [C.c_void_p, C.c_int]
        L.b2o_set_gravity.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.b2o_joint_set_motor_speed.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.b2o_joint_set_max_motor_torque.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.b2o_joint_enable_motor.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2o_joint_enable_limit.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2o_joint_set_limits.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.b2o_set_collect_dag.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.b2o_get_dag_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.b2o_get_dag_cyclic.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.b2o_step.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        L.b2o_body_count.argtypes = [C.c_void_p]
        L.b2o_contact_count.argtypes = [C.c_void_p]
        L.b2o_get_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.b2o_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.b2o_get_events.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_get_events.restype = C.c_int
        L.b2o_get_post_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_get_post_solve.restype = C.c_int
        L.b2o_ray_cast_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_query_aabb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.b2o_snapshot_sizes.argtypes = [C.c_void_p, C.POINTER(abi.SnapshotSizes)]
        L.b2o_snapshot_export.argtypes = [C.c_void_p, C.POINTER(abi.SnapshotC)]
        L.b2o_get_body_state.argtypes = [C.c_void_p, C.c_void_p]
        L.b2o_run_worlds_mt.restype = C.c_double
        L.b2o_run_worlds_mt.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.b2o_polygon_set_as_box.argtypes = [C.POINTER(abi.ShapeDef), C.c_float, C.c_float]
        L.b2o_polygon_set_as_box_angle.argtypes = [C.POINTER(abi.ShapeDef), C.c_float, C.c_float, C.c_float, C.c_float,
                                                   C.c_float]
        L.b2o_polygon_set.argtypes = [C.POINTER(abi.ShapeDef), C.POINTER(C.c_float), C.c_int]
        L.b2o_shape_compute_mass.argtypes = [C.POINTER(abi.ShapeDef), C.c_float, C.POINTER(abi.MassData)]
        L.b2o_sweep_get_transform.argtypes = [C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
        L.b2o_sincosf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2o_shape_distance.argtypes = [C.POINTER(abi.ShapeDef), C.c_int, C.c_void_p, C.POINTER(abi.ShapeDef), C.c_int,
                                         C.c_void_p, C.c_int, C.c_void_p]
        L.b2o_test_overlap_shapes.argtypes = [C.POINTER(abi.ShapeDef), C.c_int, C.c_void_p, C.POINTER(abi.ShapeDef), C.c_int,
                                              C.c_void_p]
        L.b2o_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


class Shapes:
    """Shape factory with the reference's method names (B2polygonShape::set_as_box, ::set, ...)."""

    @staticmethod
    def polygon_box(hx, hy, center=None, angle=0.0):
        s = abi.ShapeDef()
        if center is None:
            lib().b2o_polygon_set_as_box(C.byref(s), hx, hy)
        else:
            lib().b2o_polygon_set_as_box_angle(C.byref(s), hx, hy, center[0], center[1], angle)
        return s

    @staticmethod
    def polygon(vertices):
        s = abi.ShapeDef()
        flat = (C.c_float * (2 * len(vertices)))(*[c for v in vertices for c in v])
        rc = lib().b2o_polygon_set(C.byref(s), flat, len(vertices))
        if rc != 0:
            raise ValueError("degenerate polygon")
        return s

    circle = staticmethod(abi.circle_shape)
    edge_two_sided = staticmethod(abi.edge_two_sided)
    edge_one_sided = staticmethod(abi.edge_one_sided)
    chain = staticmethod(abi.chain_shape)

    @staticmethod
    def compute_mass(shape, density):
        md = abi.MassData()
        lib().b2o_shape_compute_mass(C.byref(shape), density, C.byref(md))
        return md


class B2body:
    def __init__(self, world, index):
        self.world, self.index = world, index

    def create_fixture(self, fixture_def, shape):
        return lib().b2o_create_fixture(self.world.h, self.index, C.byref(fixture_def), C.byref(shape))

    def create_fixture_by_shape(self, shape, density):
        return self.create_fixture(abi.FixtureDef(density=density), shape)

    def set_transform(self, position, angle):
        lib().b2o_set_transform(self.world.h, self.index, position[0], position[1], angle)

    def set_linear_velocity(self, v):
        lib().b2o_set_linear_velocity(self.world.h, self.index, v[0], v[1])

    def set_angular_velocity(self, w):
        lib().b2o_set_angular_velocity(self.world.h, self.index, w)

    def apply_force_to_center(self, f, wake=True):
        lib().b2o_apply_force_to_center(self.world.h, self.index, f[0], f[1], int(wake))

    def apply_force(self, f, point, wake=True):
        lib().b2o_apply_force(self.world.h, self.index, f[0], f[1], point[0], point[1], int(wake))

    def apply_torque(self, torque, wake=True):
        lib().b2o_apply_torque(self.world.h, self.index, torque, int(wake))

    def apply_linear_impulse(self, impulse, point, wake=True):
        lib().b2o_apply_linear_impulse(self.world.h, self.index, impulse[0], impulse[1], point[0], point[1], int(wake))

    def apply_linear_impulse_to_center(self, impulse, wake=True):
        lib().b2o_apply_linear_impulse_to_center(self.world.h, self.index, impulse[0], impulse[1], int(wake))

    def apply_angular_impulse(self, impulse, wake=True):
        lib().b2o_apply_angular_impulse(self.world.h, self.index, impulse, int(wake))

    def set_awake(self, flag):
        lib().b2o_body_set_awake(self.world.h, self.index, int(flag))

    def set_damping(self, linear_damping, angular_damping):
        lib().b2o_body_set_damping(self.world.h, self.index, linear_damping, angular_damping)

    def set_gravity_scale(self, scale):
        lib().b2o_body_set_gravity_scale(self.world.h, self.index, scale)

    def set_sleeping_allowed(self, flag):
        lib().b2o_body_set_sleeping_allowed(self.world.h, self.index, int(flag))

    def _rec(self):
        return self.world.snapshot().bodies[self.index]

    def get_position(self):
        r = self._rec()
        return float(r["xf"][0]), float(r["xf"][1])

    def get_angle(self):
        return float(self._rec()["a"])


class B2joint:
    """B2revoluteJoint / B2distanceJoint handle (setters of src/joints/b2_revolute_joint.rs:172-242)."""

    def __init__(self, world, index):
        self.world, self.index = world, index

    def set_motor_speed(self, speed):
        lib().b2o_joint_set_motor_speed(self.world.h, self.index, speed)

    def set_max_motor_torque(self, torque):
        lib().b2o_joint_set_max_motor_torque(self.world.h, self.index, torque)

    def enable_motor(self, flag):
        lib().b2o_joint_enable_motor(self.world.h, self.index, int(flag))

    def enable_limit(self, flag):
        lib().b2o_joint_enable_limit(self.world.h, self.index, int(flag))

    def set_limits(self, lower, upper):
        lib().b2o_joint_set_limits(self.world.h, self.index, lower, upper)

    def set_target(self, target):
        lib().b2o_joint_set_target(self.world.h, self.index, target[0], target[1])


def _body_index(b):
    return b.index if hasattr(b, "index") else int(b)


class B2world:
    shapes = Shapes

    def __init__(self, gravity, _handle=None):
        self.h = C.c_void_p(_handle if _handle is not None else lib().b2o_world_create(gravity[0], gravity[1]))
        self._joint_handles = []

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2o_world_destroy(self.h)
            self.h = None

    def clone(self):
        return B2world(None, _handle=lib().b2o_world_clone(self.h))

    def create_body(self, body_def):
        return B2body(self, lib().b2o_create_body(self.h, C.byref(body_def)))

    def body(self, index):
        return B2body(self, index)

    def revolute_joint_def(self, body_a, body_b, anchor):
        """B2revoluteJointDef::default() + initialize(body_a, body_b, anchor)."""
        d = abi.JointDef()
        lib().b2o_revolute_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), anchor[0], anchor[1])
        return d

    def distance_joint_def(self, body_a, body_b, anchor_a, anchor_b):
        """B2distanceJointDef::default() + initialize(b1, b2, anchor1, anchor2)."""
        d = abi.JointDef()
        lib().b2o_distance_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), anchor_a[0], anchor_a[1],
                                     anchor_b[0], anchor_b[1])
        return d

    def prismatic_joint_def(self, body_a, body_b, anchor, axis):
        """B2prismaticJointDef::default() + initialize(body_a, body_b, anchor, axis)."""
        d = abi.JointDef()
        lib().b2o_prismatic_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), anchor[0], anchor[1],
                                      axis[0], axis[1])
        return d

    def friction_joint_def(self, body_a, body_b, anchor):
        """B2frictionJointDef::default() + initialize(body_a, body_b, anchor): set length (= max_force), max_motor_torque."""
        d = abi.JointDef()
        lib().b2o_friction_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), anchor[0], anchor[1])
        return d

    def motor_joint_def(self, body_a, body_b):
        """B2motorJointDef::default() + initialize(body_a, body_b)."""
        d = abi.JointDef()
        lib().b2o_motor_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b))
        return d

    def pulley_joint_def(self, body_a, body_b, ground_a, ground_b, anchor_a, anchor_b, ratio):
        """B2pulleyJointDef::default() + initialize(body_a, body_b, ground_a, ground_b, anchor_a, anchor_b, ratio)."""
        d = abi.JointDef()
        lib().b2o_pulley_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), ground_a[0], ground_a[1],
                                   ground_b[0], ground_b[1], anchor_a[0], anchor_a[1], anchor_b[0], anchor_b[1], ratio)
        return d

    def gear_joint_def(self, joint1, joint2, ratio):
        """B2gearJointDef::default() with joint1, joint2 (revolute / prismatic handles) and ratio."""
        d = abi.JointDef()
        lib().b2o_gear_joint_def(self.h, C.byref(d), joint1.index, joint2.index, ratio)
        return d

    def mouse_joint_def(self, body_a, body_b, target):
        """B2mouseJointDef::default() with `target`: set length (= max_force), stiffness, damping."""
        d = abi.JointDef()
        lib().b2o_mouse_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), target[0], target[1])
        return d

    def wheel_joint_def(self, body_a, body_b, anchor, axis):
        """B2wheelJointDef::default() + initialize(body_a, body_b, anchor, axis)."""
        d = abi.JointDef()
        lib().b2o_wheel_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), anchor[0], anchor[1],
                                  axis[0], axis[1])
        return d

    def weld_joint_def(self, body_a, body_b, anchor):
        """B2weldJointDef::default() + initialize(body_a, body_b, anchor)."""
        d = abi.JointDef()
        lib().b2o_weld_joint_def(self.h, C.byref(d), _body_index(body_a), _body_index(body_b), anchor[0], anchor[1])
        return d

    def angular_stiffness(self, frequency_hertz, damping_ratio, body_a, body_b):
        """b2_angular_stiffness: (stiffness, damping)."""
        k, d = C.c_float(), C.c_float()
        lib().b2o_angular_stiffness(self.h, frequency_hertz, damping_ratio, _body_index(body_a), _body_index(body_b),
                                    C.byref(k), C.byref(d))
        return k.value, d.value

    def linear_stiffness(self, frequency_hertz, damping_ratio, body_a, body_b):
        """b2_linear_stiffness: (stiffness, damping)."""
        k, d = C.c_float(), C.c_float()
        lib().b2o_linear_stiffness(self.h, frequency_hertz, damping_ratio, _body_index(body_a), _body_index(body_b),
                                   C.byref(k), C.byref(d))
        return k.value, d.value

    def create_joint(self, joint_def):
        j = B2joint(self, lib().b2o_create_joint(self.h, C.byref(joint_def)))
        self._joint_handles.append(j)
        return j

    def destroy_joint(self, joint):
        """B2world::destroy_joint: later joints move down one index (live handles follow)."""
        i = joint.index
        lib().b2o_destroy_joint(self.h, i)
        self._joint_handles = [h for h in self._joint_handles if h is not joint]
        joint.index = -1
        for h in self._joint_handles:
            if h.index > i:
                h.index -= 1

    def joint(self, index):
        return B2joint(self, index)

    def get_joint_count(self):
        return lib().b2o_joint_count(self.h)

    def set_allow_sleeping(self, flag):
        lib().b2o_set_allow_sleeping(self.h, int(flag))

    def set_gravity(self, gravity):
        lib().b2o_set_gravity(self.h, gravity[0], gravity[1])

    def set_warm_starting(self, flag):
        lib().b2o_set_warm_starting(self.h, int(flag))

    def set_continuous_physics(self, flag):
        if flag:
            raise NotImplementedError("TOI sub-stepping is out of scope (BASELINE.json north_star)")

    def set_block_solve(self, flag):
        lib().b2o_set_block_solve(self.h, int(flag))

    def set_collect_levels(self, flag):
        lib().b2o_set_collect_levels(self.h, int(flag))

    def set_collect_dag(self, flag, handover=2.0):
        lib().b2o_set_collect_dag(self.h, int(flag), float(handover))

    def dag_stats(self):
        """Largest island of the last step: dependency-DAG depth of the exact-order velocity sweeps and the simulated
        makespan (in visits) of the chunked dataflow schedule for 256 / 1024 / 4096 / 16384 workers."""
        out = (C.c_double * 10)()
        lib().b2o_get_dag_stats(self.h, out)
        keys = ("contacts", "bodies", "sweeps", "depth", "depth_one_sweep", "handover", "makespan_256", "makespan_1024",
                "makespan_4096", "makespan_16384")
        d = dict(zip(keys, list(out)))
        cyc = (C.c_double * 12)()
        lib().b2o_get_dag_cyclic(self.h, cyc)
        d["cyclic"] = {"chunk%d_workers%d" % (c, p): cyc[3 * i + j] for i, c in enumerate((1, 4, 16, 64))
                       for j, p in enumerate((2048, 16384, 65536))}
        return d

    def step(self, dt, velocity_iterations, position_iterations):
        lib().b2o_step(self.h, dt, velocity_iterations, position_iterations)

    def get_body_count(self):
        return lib().b2o_body_count(self.h)

    def get_contact_count(self):
        return lib().b2o_contact_count(self.h)

    def get_profile(self):
        out = (C.c_double * 7)()
        lib().b2o_get_profile(self.h, out)
        return dict(zip(("step", "collide", "solve", "solve_init", "solve_velocity", "solve_position", "broadphase"),
                        list(out)))

    def get_stats(self):
        out = np.zeros(1, abi.STATS_DTYPE)
        lib().b2o_get_stats(self.h, out.ctypes.data)
        return out[0]

    def ray_cast_closest(self, p1p2):
        rays = np.ascontiguousarray(p1p2, np.float32).reshape(-1, 4)
        out = np.zeros(rays.shape[0], abi.RAY_HIT_DTYPE)
        lib().b2o_ray_cast_closest(self.h, rays.ctypes.data, rays.shape[0], out.ctypes.data)
        return out

    def query_aabb(self, aabbs, max_hits=64):
        boxes = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 4)
        counts = np.zeros(boxes.shape[0], np.int32)
        hits = np.zeros((boxes.shape[0], max(max_hits, 1), 2), np.int32)
        lib().b2o_query_aabb(self.h, boxes.ctypes.data, boxes.shape[0], max_hits, counts.ctypes.data, hits.ctypes.data)
        return [[(int(f), int(c)) for f, c in hits[i, :min(int(counts[i]), max_hits)]] for i in range(boxes.shape[0])], counts

    def contact_events(self):
        """begin_contact (1) / end_contact (2) events of the last step in the reference's firing order:
        int32 [n][5] = (type, fixture_a, index_a, fixture_b, index_b)."""
        n = lib().b2o_get_events(self.h, None, 0)
        out = np.zeros((max(n, 1), 5), np.int32)
        lib().b2o_get_events(self.h, out.ctypes.data, n)
        return out[:n]

    def post_solve_events(self):
        """post_solve reports of the last step (B2island::report order): abi.POST_SOLVE_DTYPE array."""
        n = lib().b2o_get_post_solve(self.h, None, 0)
        out = np.zeros(max(n, 1), abi.POST_SOLVE_DTYPE)
        lib().b2o_get_post_solve(self.h, out.ctypes.data, n)
        return out[:n]

    def snapshot(self):
        n = abi.SnapshotSizes()
        lib().b2o_snapshot_sizes(self.h, C.byref(n))
        snap = abi.Snapshot(n)
        c = snap.as_c()
        rc = lib().b2o_snapshot_export(self.h, C.byref(c))
        assert rc == 0
        return snap.finish(c)

    def body_state(self):
        out = np.zeros((self.get_body_count(), 8), np.float32)
        lib().b2o_get_body_state(self.h, out.ctypes.data)
        return out


def run_worlds_mt(worlds, steps, dt, vi, pi, threads):
    """CPU baseline: one world per host thread. Returns wall seconds."""
    arr = (C.c_void_p * len(worlds))(*[w.h for w in worlds])
    return lib().b2o_run_worlds_mt(arr, len(worlds), steps, dt, vi, pi, threads)


def sincosf(angles):
    """glibc sinf/cosf of a float32 array (reference for the device trigonometry)."""
    a = np.ascontiguousarray(angles, np.float32)
    s = np.empty_like(a)
    c = np.empty_like(a)
    lib().b2o_sincosf(a.ctypes.data, s.ctypes.data, c.ctypes.data, a.size)
    return s, c


def hardware_threads():
    return int(lib().b2o_hardware_threads())


def shape_distance(shape_a, xf_a, shape_b, xf_b, use_radii=True, index_a=0, index_b=0):
    """b2_distance_fn (GJK) between two shapes under (x, y, angle) transforms: (point_a, point_b, distance, iterations)."""
    import numpy as np
    xa = np.asarray(xf_a, np.float32)
    xb = np.asarray(xf_b, np.float32)
    out = np.zeros(5, np.float32)
    it = lib().b2o_shape_distance(C.byref(shape_a), index_a, xa.ctypes.data, C.byref(shape_b), index_b, xb.ctypes.data,
                                  1 if use_radii else 0, out.ctypes.data)
    return (float(out[0]), float(out[1])), (float(out[2]), float(out[3])), float(out[4]), it


def test_overlap_shapes(shape_a, xf_a, shape_b, xf_b, index_a=0, index_b=0):
    import numpy as np
    xa = np.asarray(xf_a, np.float32)
    xb = np.asarray(xf_b, np.float32)
    return bool(lib().b2o_test_overlap_shapes(C.byref(shape_a), index_a, xa.ctypes.data, C.byref(shape_b), index_b, xb.ctypes.data))
