"""ctypes / numpy mirror of include/b2gpu.h (the C ABI of the step engine).

Record layouts are asserted against the sizes documented in the header so a drift
between the header and this file fails at import time.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 2

OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_CAPACITY, E_UNSUPPORTED, E_LOCKED, E_INTERNAL, E_IO = 0, -1, -2, -3, -4, -5, -6, -7, -8

STATIC_BODY, KINEMATIC_BODY, DYNAMIC_BODY = 0, 1, 2
BODY_ISLAND, BODY_AWAKE, BODY_AUTO_SLEEP, BODY_BULLET, BODY_FIXED_ROTATION, BODY_ENABLED, BODY_TOI = (
    0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40)
CONTACT_ISLAND, CONTACT_TOUCHING, CONTACT_ENABLED, CONTACT_FILTER = 0x1, 0x2, 0x4, 0x8
SHAPE_CIRCLE, SHAPE_EDGE, SHAPE_POLYGON, SHAPE_CHAIN = 0, 1, 2, 3
MANIFOLD_CIRCLES, MANIFOLD_FACE_A, MANIFOLD_FACE_B = 0, 1, 2
WORLD_ALLOW_SLEEP, WORLD_WARM_STARTING, WORLD_NEW_CONTACTS, WORLD_CLEAR_FORCES, WORLD_BLOCK_SOLVE = (
    0x01, 0x02, 0x04, 0x08, 0x10)
MAX_POLYGON_VERTICES = 8
JOINT_DISTANCE, JOINT_FRICTION, JOINT_MOTOR, JOINT_PRISMATIC, JOINT_REVOLUTE, JOINT_WELD, JOINT_WHEEL = 1, 2, 4, 6, 8, 9, 10
JOINT_GEAR, JOINT_MOUSE, JOINT_PULLEY = 3, 5, 7
JOINT_CONTROL_MOTOR_SPEED, JOINT_CONTROL_MAX_MOTOR_TORQUE, JOINT_CONTROL_TARGET = 0, 1, 2  # b2gpu_batch_set_joint_control  # B2jointType (src/b2_joint.rs:46-58)
JOINT_COLLIDE_CONNECTED, JOINT_ENABLE_LIMIT, JOINT_ENABLE_MOTOR = 0x1, 0x2, 0x4
POLYGON_RADIUS = float(np.float32(2.0) * np.float32(0.005))  # src/b2_common.rs:48

f32, i32, u32, u16, i16 = np.float32, np.int32, np.uint32, np.uint16, np.int16

BODY_DTYPE = np.dtype([
    ("type", i32), ("flags", u32),
    ("xf", f32, 4),          # p.x p.y q.s q.c
    ("lc", f32, 2), ("c0", f32, 2), ("c", f32, 2), ("a0", f32), ("a", f32),
    ("v", f32, 2), ("w", f32),
    ("force", f32, 2), ("torque", f32),
    ("mass", f32), ("inv_mass", f32), ("inertia", f32), ("inv_inertia", f32),
    ("linear_damping", f32), ("angular_damping", f32), ("gravity_scale", f32), ("sleep_time", f32),
    ("fixture_head", i32), ("fixture_count", i32), ("reserved", i32, 2),
], align=True)
FIXTURE_DTYPE = np.dtype([
    ("body", i32), ("next", i32), ("shape_type", i32), ("shape_first", i32), ("child_count", i32), ("proxy_first", i32),
    ("density", f32), ("friction", f32), ("restitution", f32), ("restitution_threshold", f32),
    ("category_bits", u16), ("mask_bits", u16), ("group_index", i16), ("is_sensor", u16),
], align=True)
SHAPE_DTYPE = np.dtype([
    ("type", i32), ("radius", f32), ("count", i32), ("one_sided", i32), ("c", f32, 2),
    ("v", f32, 16), ("n", f32, 16), ("reserved", i32, 2),
], align=True)
PROXY_DTYPE = np.dtype([
    ("fixture", i32), ("child_index", i32), ("proxy_id", i32), ("reserved", i32), ("aabb", f32, 4),
], align=True)
NODE_DTYPE = np.dtype([
    ("aabb", f32, 4), ("parent", i32), ("child1", i32), ("child2", i32), ("height", i32), ("proxy", i32), ("moved", i32),
], align=True)
MPOINT_DTYPE = np.dtype([
    ("lp", f32, 2), ("normal_impulse", f32), ("tangent_impulse", f32), ("id", u32),
], align=True)
MANIFOLD_DTYPE = np.dtype([
    ("points", MPOINT_DTYPE, 2), ("ln", f32, 2), ("lp", f32, 2), ("type", i32), ("point_count", i32),
], align=True)
CONTACT_DTYPE = np.dtype([
    ("fixture_a", i32), ("fixture_b", i32), ("index_a", i32), ("index_b", i32), ("flags", u32),
    ("friction", f32), ("restitution", f32), ("restitution_threshold", f32), ("tangent_speed", f32), ("reserved", i32),
    ("manifold", MANIFOLD_DTYPE),
], align=True)
JOINT_DTYPE = np.dtype([
    ("type", i32), ("body_a", i32), ("body_b", i32), ("flags", u32),
    ("local_anchor_a", f32, 2), ("local_anchor_b", f32, 2), ("param", f32, 8), ("impulse", f32, 8),
], align=True)
STATS_DTYPE = np.dtype([
    ("status", i32), ("contacts", i32), ("touching", i32), ("destroyed", i32), ("islands", i32), ("island_bodies", i32),
    ("island_contacts", i32), ("moved", i32), ("pairs", i32), ("created", i32), ("awake_bodies", i32),
    ("solver_levels", i32), ("reserved", i32, 4),
], align=True)

assert BODY_DTYPE.itemsize == 128, BODY_DTYPE.itemsize
assert FIXTURE_DTYPE.itemsize == 48, FIXTURE_DTYPE.itemsize
assert SHAPE_DTYPE.itemsize == 160, SHAPE_DTYPE.itemsize
assert PROXY_DTYPE.itemsize == 32
assert NODE_DTYPE.itemsize == 40
assert MANIFOLD_DTYPE.itemsize == 64, MANIFOLD_DTYPE.itemsize
assert CONTACT_DTYPE.itemsize == 104, CONTACT_DTYPE.itemsize
assert STATS_DTYPE.itemsize == 64
assert JOINT_DTYPE.itemsize == 96, JOINT_DTYPE.itemsize


class WorldRec(C.Structure):
    _fields_ = [("gravity_x", C.c_float), ("gravity_y", C.c_float), ("inv_dt0", C.c_float), ("flags", C.c_uint32),
                ("tree_root", C.c_int32), ("tree_free_list", C.c_int32), ("tree_node_count", C.c_int32),
                ("tree_node_capacity", C.c_int32), ("tree_insertion_count", C.c_int32), ("proxy_count", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class SnapshotSizes(C.Structure):
    _fields_ = [("body_count", C.c_int32), ("fixture_count", C.c_int32), ("shape_count", C.c_int32),
                ("proxy_count", C.c_int32), ("node_count", C.c_int32), ("contact_count", C.c_int32),
                ("move_count", C.c_int32), ("joint_count", C.c_int32)]


class SnapshotC(C.Structure):
    _fields_ = [("world", WorldRec), ("n", SnapshotSizes), ("bodies", C.c_void_p), ("fixtures", C.c_void_p),
                ("shapes", C.c_void_p), ("proxies", C.c_void_p), ("nodes", C.c_void_p), ("contacts", C.c_void_p),
                ("move_buffer", C.c_void_p), ("joints", C.c_void_p)]


class BodyDef(C.Structure):
    """B2bodyDef (src/b2_body.rs:39-58) with the reference's defaults."""
    _fields_ = [("type", C.c_int32), ("position_x", C.c_float), ("position_y", C.c_float), ("angle", C.c_float),
                ("linear_velocity_x", C.c_float), ("linear_velocity_y", C.c_float), ("angular_velocity", C.c_float),
                ("linear_damping", C.c_float), ("angular_damping", C.c_float),
                ("allow_sleep", C.c_int32), ("awake", C.c_int32), ("fixed_rotation", C.c_int32),
                ("bullet", C.c_int32), ("enabled", C.c_int32), ("gravity_scale", C.c_float)]

    def __init__(self, **kw):
        super().__init__()
        self.type = STATIC_BODY
        self.allow_sleep = 1
        self.awake = 1
        self.enabled = 1
        self.gravity_scale = 1.0
        for k, v in kw.items():
            if k == "position":
                self.position_x, self.position_y = v
            elif k == "linear_velocity":
                self.linear_velocity_x, self.linear_velocity_y = v
            else:
                setattr(self, k, v)


class FixtureDef(C.Structure):
    """B2fixtureDef (src/b2_fixture.rs:44-57) with the reference's defaults."""
    _fields_ = [("friction", C.c_float), ("restitution", C.c_float), ("restitution_threshold", C.c_float),
                ("density", C.c_float), ("is_sensor", C.c_int32), ("category_bits", C.c_uint16),
                ("mask_bits", C.c_uint16), ("group_index", C.c_int16), ("reserved", C.c_uint16)]

    def __init__(self, **kw):
        super().__init__()
        self.friction = 0.2
        self.restitution = 0.0
        self.restitution_threshold = 1.0
        self.density = 0.0
        self.category_bits = 0x0001
        self.mask_bits = 0xFFFF
        self.group_index = 0
        for k, v in kw.items():
            setattr(self, k, v)


class ShapeDef(C.Structure):
    _fields_ = [("type", C.c_int32), ("radius", C.c_float), ("p_x", C.c_float), ("p_y", C.c_float),
                ("v0", C.c_float * 2), ("v1", C.c_float * 2), ("v2", C.c_float * 2), ("v3", C.c_float * 2),
                ("one_sided", C.c_int32), ("count", C.c_int32), ("centroid", C.c_float * 2),
                ("vertices", C.c_float * 16), ("normals", C.c_float * 16),
                ("chain_vertices", C.POINTER(C.c_float)), ("chain_count", C.c_int32),
                ("chain_prev", C.c_float * 2), ("chain_next", C.c_float * 2)]


class JointDef(C.Structure):
    """b2gpu_joint_def: B2revoluteJointDef / B2distanceJointDef as one plain struct (filled by
    world.revolute_joint_def / world.distance_joint_def = the reference's Default + initialize)."""
    _fields_ = [("type", C.c_int32), ("body_a", C.c_int32), ("body_b", C.c_int32), ("collide_connected", C.c_int32),
                ("local_anchor_a", C.c_float * 2), ("local_anchor_b", C.c_float * 2),
                ("reference_angle", C.c_float), ("lower_angle", C.c_float), ("upper_angle", C.c_float),
                ("max_motor_torque", C.c_float), ("motor_speed", C.c_float),
                ("enable_limit", C.c_int32), ("enable_motor", C.c_int32),
                ("length", C.c_float), ("min_length", C.c_float), ("max_length", C.c_float),
                ("stiffness", C.c_float), ("damping", C.c_float)]


class MassData(C.Structure):
    _fields_ = [("mass", C.c_float), ("center_x", C.c_float), ("center_y", C.c_float), ("inertia", C.c_float)]


class Caps(C.Structure):
    _fields_ = [("max_bodies", C.c_int32), ("max_fixtures", C.c_int32), ("max_shapes", C.c_int32),
                ("max_proxies", C.c_int32), ("max_contacts", C.c_int32), ("max_pairs", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class Snapshot:
    """Full step state as numpy record arrays (caller-owned side of b2gpu_snapshot)."""

    def __init__(self, sizes):
        self.world = WorldRec()
        self.alloc(sizes)

    def alloc(self, n):
        self.bodies = np.zeros(max(n.body_count, 1), BODY_DTYPE)
        self.fixtures = np.zeros(max(n.fixture_count, 1), FIXTURE_DTYPE)
        self.shapes = np.zeros(max(n.shape_count, 1), SHAPE_DTYPE)
        self.proxies = np.zeros(max(n.proxy_count, 1), PROXY_DTYPE)
        self.nodes = np.zeros(max(n.node_count, 1), NODE_DTYPE)
        self.contacts = np.zeros(max(n.contact_count, 1), CONTACT_DTYPE)
        self.move_buffer = np.zeros(max(n.move_count, 1), np.int32)
        self.joints = np.zeros(max(n.joint_count, 1), JOINT_DTYPE)
        self.n = SnapshotSizes(n.body_count, n.fixture_count, n.shape_count, n.proxy_count, n.node_count,
                               n.contact_count, n.move_count, n.joint_count)

    def as_c(self):
        s = SnapshotC()
        s.world = self.world
        s.n = self.n
        s.bodies = self.bodies.ctypes.data
        s.fixtures = self.fixtures.ctypes.data
        s.shapes = self.shapes.ctypes.data
        s.proxies = self.proxies.ctypes.data
        s.nodes = self.nodes.ctypes.data
        s.contacts = self.contacts.ctypes.data
        s.move_buffer = self.move_buffer.ctypes.data
        s.joints = self.joints.ctypes.data
        return s

    def finish(self, c):
        """Adopt counts/world scalars written by an export/download call and trim the arrays."""
        self.world = c.world
        self.n = c.n
        n = c.n
        self.bodies = self.bodies[:n.body_count]
        self.fixtures = self.fixtures[:n.fixture_count]
        self.shapes = self.shapes[:n.shape_count]
        self.proxies = self.proxies[:n.proxy_count]
        self.nodes = self.nodes[:n.node_count]
        self.contacts = self.contacts[:n.contact_count]
        self.move_buffer = self.move_buffer[:n.move_count]
        self.joints = self.joints[:n.joint_count]
        return self


# ---- shape helpers shared by both mirrors (they only fill plain ShapeDef fields)
def circle_shape(radius, p=(0.0, 0.0)):
    s = ShapeDef()
    s.type = SHAPE_CIRCLE
    s.radius = radius
    s.p_x, s.p_y = p
    return s


def edge_two_sided(v1, v2):
    """B2edgeShape::set_two_sided (b2_edge_shape.rs(private):15-19); radius = B2_POLYGON_RADIUS."""
    s = ShapeDef()
    s.type = SHAPE_EDGE
    s.radius = POLYGON_RADIUS
    s.v1[0], s.v1[1] = v1
    s.v2[0], s.v2[1] = v2
    s.one_sided = 0
    return s


def edge_one_sided(v0, v1, v2, v3):
    s = ShapeDef()
    s.type = SHAPE_EDGE
    s.radius = POLYGON_RADIUS
    for dst, src in ((s.v0, v0), (s.v1, v1), (s.v2, v2), (s.v3, v3)):
        dst[0], dst[1] = src
    s.one_sided = 1
    return s


def chain_shape(vertices, prev_vertex, next_vertex, loop=False):
    """B2chainShape::create_chain / create_loop (b2_chain_shape.rs(private):12-46)."""
    vs = [tuple(v) for v in vertices]
    if loop:
        vs = vs + [vs[0]]
        prev_vertex, next_vertex = vs[-2], vs[1]
    s = ShapeDef()
    s.type = SHAPE_CHAIN
    s.radius = POLYGON_RADIUS
    arr = (C.c_float * (2 * len(vs)))(*[c for v in vs for c in v])
    s._keep = arr  # keep the buffer alive
    s.chain_vertices = C.cast(arr, C.POINTER(C.c_float))
    s.chain_count = len(vs)
    s.chain_prev[0], s.chain_prev[1] = prev_vertex
    s.chain_next[0], s.chain_next[1] = next_vertex
    return s


# b2gpu_contact_event (include/b2gpu.h)
EVENT_BEGIN_CONTACT, EVENT_END_CONTACT = 1, 2
CONTACT_EVENT_DTYPE = np.dtype([("type", np.int32), ("fixture_a", np.int32), ("index_a", np.int32), ("fixture_b", np.int32),
                                ("index_b", np.int32), ("reserved", np.int32, 3)])
assert CONTACT_EVENT_DTYPE.itemsize == 32

# b2gpu_ray_hit (include/b2gpu.h)
RAY_HIT_DTYPE = np.dtype([("fixture", np.int32), ("child_index", np.int32), ("fraction", np.float32), ("point", np.float32, 2),
                          ("normal", np.float32, 2), ("reserved", np.int32)])

# b2gpu_post_solve_event (include/b2gpu.h)
POST_SOLVE_DTYPE = np.dtype([("fixture_a", np.int32), ("index_a", np.int32), ("fixture_b", np.int32), ("index_b", np.int32),
                             ("count", np.int32), ("normal_impulses", np.float32, 2), ("tangent_impulses", np.float32, 2),
                             ("reserved", np.int32, 3)])
assert POST_SOLVE_DTYPE.itemsize == 48
