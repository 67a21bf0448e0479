"""World DEFINITIONS in the layout of the crate's serde support (feature `serde_support`,
src/serialize/serialize_b2_world.rs:133-178, serialize_b2_body.rs:36-70, serialize_b2_fixture.rs:23-52,
serialize_b2_joint.rs:33-53, src/joints/serialize/*.rs, src/shapes/serialize_b2_polygon_shape.rs:40-95).

`to_serde(snapshot)` turns a b2gpu snapshot into the document `serde_json::to_string(&world)` would write for the same
world — field for field, in the crate's order conventions:

* `m_bodies_list` iterates the world's body list, NEWEST body first; a joint names its bodies by their position in that list
  (the serializer stores it in `m_island_index`);
* every body carries `m_definition` (`B2bodyDef` read back from the live body: position = m_xf.p, angle = m_sweep.a, the flag
  bits) and `m_fixture_list`, newest fixture first, each with its shape struct and `m_shape_type`;
* `m_joints_list` holds `{jtype, joint_def}` of every joint but gear and mouse joints, newest first; mouse joints are skipped
  (the crate skips them too) but still count in the numbering gear joints use for `joint1` / `joint2`;
  `m_gear_joints_list` follows.

`from_serde(doc, world)` rebuilds a world through the public API (oracle mirror or device mirror) in the order the crate's
`Deserialize` does: bodies in document order — so, as in the crate, a save / load round trip reverses the creation order, and a
second round trip restores it.  Like the crate's files this is a DEFINITION: contacts, warm-start impulses and the broadphase
tree are not part of it (the checkpoint format of b2gpu_snapshot_save is what resumes a run).

Written from the crate's Serialize / Deserialize impls; not checked against files produced by the crate (no Rust toolchain in
this image) — tests/test_serde_world.py checks the layout against the field lists above and the double round trip."""
import numpy as np

from . import abi

BODY_TYPES = ["B2StaticBody", "B2KinematicBody", "B2DynamicBody"]          # src/b2_body.rs:27-31
SHAPE_TYPES = {abi.SHAPE_CIRCLE: "ECircle", abi.SHAPE_EDGE: "EEdge", abi.SHAPE_POLYGON: "EPolygon", abi.SHAPE_CHAIN: "EChain"}
JOINT_TYPES = {abi.JOINT_DISTANCE: "EDistanceJoint", abi.JOINT_FRICTION: "EFrictionJoint", abi.JOINT_GEAR: "EGearJoint",
               abi.JOINT_MOTOR: "EMotorJoint", abi.JOINT_MOUSE: "EMouseJoint", abi.JOINT_PRISMATIC: "EPrismaticJoint",
               abi.JOINT_PULLEY: "EPulleyJoint", abi.JOINT_REVOLUTE: "ERevoluteJoint", abi.JOINT_WELD: "EWeldJoint",
               abi.JOINT_WHEEL: "EWheelJoint"}                              # src/b2_joint.rs:46-58
JOINT_CODES = {v: k for k, v in JOINT_TYPES.items()}
SHAPE_CODES = {v: k for k, v in SHAPE_TYPES.items()}


def _f(x):
    return float(np.float32(x))


def _vec(x, y):
    return {"x": _f(x), "y": _f(y)}


def _shape(snap, fx):
    t = int(fx["shape_type"])
    first, n = int(fx["shape_first"]), int(fx["child_count"])
    s0 = snap.shapes[first]
    base = {"m_type": SHAPE_TYPES[t], "m_radius": _f(s0["radius"])}
    if t == abi.SHAPE_CIRCLE:
        return {"base": base, "m_p": _vec(*s0["c"])}
    if t == abi.SHAPE_EDGE:
        v = s0["v"]
        return {"base": base, "m_vertex1": _vec(v[2], v[3]), "m_vertex2": _vec(v[4], v[5]), "m_vertex0": _vec(v[0], v[1]),
                "m_vertex3": _vec(v[6], v[7]), "m_one_sided": bool(s0["one_sided"])}
    if t == abi.SHAPE_POLYGON:
        c = int(s0["count"])
        return {"base": base, "m_centroid": _vec(*s0["c"]), "m_count": c,
                "m_vertices": [_vec(s0["v"][2 * i], s0["v"][2 * i + 1]) for i in range(c)],
                "m_normals": [_vec(s0["n"][2 * i], s0["n"][2 * i + 1]) for i in range(c)]}
    # chain: children are its edges (v0 = previous ghost, v1, v2, v3 = next ghost); the vertex list is v1 of the first + every v2
    kids = [snap.shapes[first + i]["v"] for i in range(n)]
    verts = [_vec(kids[0][2], kids[0][3])] + [_vec(k[4], k[5]) for k in kids]
    return {"base": base, "m_vertices": verts, "m_prev_vertex": _vec(kids[0][0], kids[0][1]),
            "m_next_vertex": _vec(kids[-1][6], kids[-1][7])}


def _joint_def(j, body_pos, joint_pos, find_coupled):
    t = int(j["type"])
    p, fl = j["param"], int(j["flags"])
    base = {"jtype": JOINT_TYPES[t], "user_data": None, "body_a": body_pos[int(j["body_a"])], "body_b": body_pos[int(j["body_b"])],
            "collide_connected": bool(fl & abi.JOINT_COLLIDE_CONNECTED)}
    la, lb = _vec(*j["local_anchor_a"]), _vec(*j["local_anchor_b"])
    lim, mot = bool(fl & abi.JOINT_ENABLE_LIMIT), bool(fl & abi.JOINT_ENABLE_MOTOR)
    if t == abi.JOINT_REVOLUTE:
        return {"base": base, "local_anchor_a": la, "local_anchor_b": lb, "reference_angle": _f(p[0]), "enable_limit": lim,
                "lower_angle": _f(p[1]), "upper_angle": _f(p[2]), "enable_motor": mot, "motor_speed": _f(p[4]),
                "max_motor_torque": _f(p[3])}
    if t == abi.JOINT_PRISMATIC:
        return {"base": base, "local_anchor_a": la, "local_anchor_b": lb, "local_axis_a": _vec(p[5], p[6]),
                "reference_angle": _f(p[0]), "enable_limit": lim, "lower_translation": _f(p[1]), "upper_translation": _f(p[2]),
                "enable_motor": mot, "motor_speed": _f(p[4]), "max_motor_force": _f(p[3])}
    if t == abi.JOINT_DISTANCE:
        return {"base": base, "local_anchor_a": la, "local_anchor_b": lb, "length": _f(p[0]), "min_length": _f(p[1]),
                "max_length": _f(p[2]), "stiffness": _f(p[3]), "damping": _f(p[4])}
    if t == abi.JOINT_WELD:
        return {"base": base, "local_anchor_a": la, "local_anchor_b": lb, "reference_angle": _f(p[0]), "stiffness": _f(p[3]),
                "damping": _f(p[4])}
    if t == abi.JOINT_WHEEL:
        return {"base": base, "local_anchor_a": la, "local_anchor_b": lb, "local_axis_a": _vec(p[5], p[6]), "enable_limit": lim,
                "lower_translation": _f(p[1]), "upper_translation": _f(p[2]), "enable_motor": mot, "motor_speed": _f(p[4]),
                "max_motor_torque": _f(p[3]), "stiffness": _f(p[0]), "damping": _f(p[7])}
    if t == abi.JOINT_FRICTION:
        return {"base": base, "local_anchor_a": la, "local_anchor_b": lb, "max_force": _f(p[0]), "max_torque": _f(p[1])}
    if t == abi.JOINT_MOTOR:
        return {"base": base, "linear_offset": la, "angular_offset": _f(p[2]), "max_force": _f(p[0]), "max_torque": _f(p[1]),
                "correction_factor": _f(p[3])}
    if t == abi.JOINT_PULLEY:
        return {"base": base, "ground_anchor_a": _vec(p[0], p[1]), "ground_anchor_b": _vec(p[2], p[3]), "local_anchor_a": la,
                "local_anchor_b": lb, "length_a": _f(p[4]), "length_b": _f(p[5]), "ratio": _f(p[6])}
    if t == abi.JOINT_GEAR:
        j1, j2 = find_coupled(j)
        return {"base": base, "joint1": joint_pos[j1], "joint2": joint_pos[j2], "ratio": _f(j["impulse"][4])}
    raise ValueError("joint type %d has no serde definition" % t)


def to_serde(snap):
    """abi.Snapshot -> dict in the crate's `impl Serialize for B2world` layout (json.dumps-able)."""
    nb = len(snap.bodies)
    order = list(range(nb - 1, -1, -1))                  # the body list iterates newest first
    body_pos = {b: i for i, b in enumerate(order)}       # what the serializer writes into m_island_index
    bodies = []
    for b in order:
        r = snap.bodies[b]
        fl = int(r["flags"])
        d = {"body_type": BODY_TYPES[int(r["type"])], "position": _vec(r["xf"][0], r["xf"][1]), "angle": _f(r["a"]),
             "linear_velocity": _vec(*r["v"]), "angular_velocity": _f(r["w"]), "linear_damping": _f(r["linear_damping"]),
             "angular_damping": _f(r["angular_damping"]), "allow_sleep": bool(fl & abi.BODY_AUTO_SLEEP), "awake": bool(fl & abi.BODY_AWAKE),
             "fixed_rotation": bool(fl & abi.BODY_FIXED_ROTATION), "bullet": bool(fl & abi.BODY_BULLET),
             "enabled": bool(fl & abi.BODY_ENABLED), "user_data": None, "gravity_scale": _f(r["gravity_scale"])}
        fixtures = []
        f = int(r["fixture_head"])
        while f != -1:                                   # newest fixture first
            fx = snap.fixtures[f]
            fixtures.append({"m_friction": _f(fx["friction"]), "m_restitution": _f(fx["restitution"]),
                             "m_restitution_threshold": _f(fx["restitution_threshold"]), "m_density": _f(fx["density"]),
                             "m_is_sensor": bool(fx["is_sensor"]),
                             "m_filter": {"category_bits": int(fx["category_bits"]), "mask_bits": int(fx["mask_bits"]),
                                          "group_index": int(fx["group_index"])},
                             "m_shape_type": SHAPE_TYPES[int(fx["shape_type"])], "m_shape": _shape(snap, fx)})
            f = int(fx["next"])
        bodies.append({"m_definition": d, "m_fixture_list": fixtures})
    nj = len(snap.joints)
    jorder = list(range(nj - 1, -1, -1))                 # the joint list iterates newest first
    plain = [j for j in jorder if int(snap.joints[j]["type"]) != abi.JOINT_GEAR]
    gears = [j for j in jorder if int(snap.joints[j]["type"]) == abi.JOINT_GEAR]
    joint_pos = {j: i for i, j in enumerate(plain)}      # m_index: mouse joints count, gear joints keep -1

    def find_coupled(g):
        """The gear record keeps copies of its coupled joints' anchors, not their indices: find them again."""
        body_c, body_d = (int(x) for x in np.array(g["impulse"][5:7], np.float32).view(np.int32))
        out = []
        for body_x, anchor_x, anchor_own, ref, prismatic in (
                (body_c, g["param"][0:2], g["local_anchor_a"], g["impulse"][1], bool(int(g["flags"]) & 0x100)),
                (body_d, g["param"][2:4], g["local_anchor_b"], g["impulse"][2], bool(int(g["flags"]) & 0x200))):
            want = abi.JOINT_PRISMATIC if prismatic else abi.JOINT_REVOLUTE
            hit = [j for j in plain if int(snap.joints[j]["type"]) == want and int(snap.joints[j]["body_a"]) == body_x
                   and np.array_equal(snap.joints[j]["local_anchor_a"], anchor_x)
                   and np.array_equal(snap.joints[j]["local_anchor_b"], anchor_own) and snap.joints[j]["param"][0] == ref]
            if not hit:
                raise ValueError("gear joint: a coupled joint is no longer in the world (destroy the gear joint first)")
            out.append(hit[0])
        return out

    def defs(ids):
        return [{"jtype": JOINT_TYPES[int(snap.joints[j]["type"])], "joint_def": _joint_def(snap.joints[j], body_pos, joint_pos, find_coupled)}
                for j in ids if int(snap.joints[j]["type"]) != abi.JOINT_MOUSE]

    return {"m_gravity": _vec(snap.world.gravity_x, snap.world.gravity_y), "m_bodies_list": bodies,
            "m_joints_list": defs(plain), "m_gear_joints_list": defs(gears)}


def _shape_def(world, t, s):
    if t == "ECircle":
        return world.shapes.circle(s["base"]["m_radius"], (s["m_p"]["x"], s["m_p"]["y"]))
    d = abi.ShapeDef()
    d.radius = s["base"]["m_radius"]
    if t == "EEdge":
        d.type = abi.SHAPE_EDGE
        for dst, key in ((d.v0, "m_vertex0"), (d.v1, "m_vertex1"), (d.v2, "m_vertex2"), (d.v3, "m_vertex3")):
            dst[0], dst[1] = s[key]["x"], s[key]["y"]
        d.one_sided = int(s["m_one_sided"])
        return d
    if t == "EPolygon":  # the fields as stored: no hull recomputation, as in the crate's Deserialize
        d.type = abi.SHAPE_POLYGON
        d.count = int(s["m_count"])
        d.centroid[0], d.centroid[1] = s["m_centroid"]["x"], s["m_centroid"]["y"]
        for i in range(d.count):
            d.vertices[2 * i], d.vertices[2 * i + 1] = s["m_vertices"][i]["x"], s["m_vertices"][i]["y"]
            d.normals[2 * i], d.normals[2 * i + 1] = s["m_normals"][i]["x"], s["m_normals"][i]["y"]
        return d
    vs = [(v["x"], v["y"]) for v in s["m_vertices"]]
    return abi.chain_shape(vs, (s["m_prev_vertex"]["x"], s["m_prev_vertex"]["y"]), (s["m_next_vertex"]["x"], s["m_next_vertex"]["y"]))


def from_serde(doc, world):
    """Rebuilds the definitions of `doc` in `world` (an empty B2world of either mirror; its gravity is set from the document).
    Returns (bodies, joints) handles in creation order."""
    world.set_gravity((doc["m_gravity"]["x"], doc["m_gravity"]["y"]))
    bodies = []
    for b in doc["m_bodies_list"]:
        d = b["m_definition"]
        bd = abi.BodyDef(type=BODY_TYPES.index(d["body_type"]), position=(d["position"]["x"], d["position"]["y"]), angle=d["angle"],
                         linear_velocity=(d["linear_velocity"]["x"], d["linear_velocity"]["y"]), angular_velocity=d["angular_velocity"],
                         linear_damping=d["linear_damping"], angular_damping=d["angular_damping"], allow_sleep=int(d["allow_sleep"]),
                         awake=int(d["awake"]), fixed_rotation=int(d["fixed_rotation"]), bullet=int(d["bullet"]),
                         enabled=int(d["enabled"]), gravity_scale=d["gravity_scale"])
        body = world.create_body(bd)
        for f in b["m_fixture_list"]:
            fd = abi.FixtureDef(friction=f["m_friction"], restitution=f["m_restitution"], restitution_threshold=f["m_restitution_threshold"],
                                density=f["m_density"], is_sensor=int(f["m_is_sensor"]), category_bits=f["m_filter"]["category_bits"],
                                mask_bits=f["m_filter"]["mask_bits"], group_index=f["m_filter"]["group_index"])
            body.create_fixture(fd, _shape_def(world, f["m_shape_type"], f["m_shape"]))
        bodies.append(body)
    joints = []
    for e in doc["m_joints_list"] + doc["m_gear_joints_list"]:
        j, t = e["joint_def"], JOINT_CODES[e["jtype"]]
        jd = abi.JointDef()
        jd.type, jd.body_a, jd.body_b = t, bodies[j["base"]["body_a"]].index, bodies[j["base"]["body_b"]].index
        jd.collide_connected = int(j["base"]["collide_connected"])
        for name in ("local_anchor_a", "local_anchor_b"):
            if name in j:
                getattr(jd, name)[0], getattr(jd, name)[1] = j[name]["x"], j[name]["y"]
        if t in (abi.JOINT_REVOLUTE, abi.JOINT_PRISMATIC, abi.JOINT_WHEEL):
            pris = t != abi.JOINT_REVOLUTE
            jd.enable_limit, jd.enable_motor, jd.motor_speed = int(j["enable_limit"]), int(j["enable_motor"]), j["motor_speed"]
            jd.lower_angle, jd.upper_angle = (j["lower_translation"], j["upper_translation"]) if pris else (j["lower_angle"], j["upper_angle"])
            jd.max_motor_torque = j["max_motor_force"] if t == abi.JOINT_PRISMATIC else j["max_motor_torque"]
            if t != abi.JOINT_WHEEL:
                jd.reference_angle = j["reference_angle"]
            if pris:  # b2gpu.h: (length, min_length) carry local_axis_a
                jd.length, jd.min_length, jd.max_length = j["local_axis_a"]["x"], j["local_axis_a"]["y"], 0.0
            if t == abi.JOINT_WHEEL:
                jd.stiffness, jd.damping = j["stiffness"], j["damping"]
        elif t == abi.JOINT_DISTANCE:
            jd.length, jd.min_length, jd.max_length, jd.stiffness, jd.damping = j["length"], j["min_length"], j["max_length"], j["stiffness"], j["damping"]
        elif t == abi.JOINT_WELD:
            jd.reference_angle, jd.stiffness, jd.damping = j["reference_angle"], j["stiffness"], j["damping"]
        elif t == abi.JOINT_FRICTION:
            jd.length, jd.max_motor_torque = j["max_force"], j["max_torque"]
        elif t == abi.JOINT_MOTOR:
            jd.local_anchor_a[0], jd.local_anchor_a[1] = j["linear_offset"]["x"], j["linear_offset"]["y"]
            jd.reference_angle, jd.length, jd.max_motor_torque, jd.stiffness = j["angular_offset"], j["max_force"], j["max_torque"], j["correction_factor"]
        elif t == abi.JOINT_PULLEY:
            jd.lower_angle, jd.upper_angle = j["ground_anchor_a"]["x"], j["ground_anchor_a"]["y"]
            jd.max_motor_torque, jd.motor_speed = j["ground_anchor_b"]["x"], j["ground_anchor_b"]["y"]
            jd.length, jd.min_length, jd.max_length = j["length_a"], j["length_b"], j["ratio"]
        elif t == abi.JOINT_GEAR:  # joint1 / joint2 count over the non-gear joints in document order (mouse joints included there)
            jd.enable_limit, jd.enable_motor, jd.length = joints[j["joint1"]].index, joints[j["joint2"]].index, j["ratio"]
        joints.append(world.create_joint(jd))
    return bodies, joints
