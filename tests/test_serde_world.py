"""box2d_rs_b200/serde_world.py: world definitions in the layout of the crate's serde support
(src/serialize/serialize_b2_world.rs:133-178 and the per-type Serialize impls).  No file written by the crate exists here (no
Rust toolchain), so the checks are: (1) the document has exactly the fields the crate's impls name, (2) it survives JSON,
(3) loading it twice restores the as-built world bit for bit — one load reverses the creation order, as the crate's own
Deserialize does —, (4) both mirrors build the same world from the same document."""
import json

import numpy as np
import pytest

import parity
from conftest import HOSTSIM_SO, JOINT_SCENES, SCENES

BODY_DEF = {"body_type", "position", "angle", "linear_velocity", "angular_velocity", "linear_damping", "angular_damping",
            "allow_sleep", "awake", "fixed_rotation", "bullet", "enabled", "user_data", "gravity_scale"}      # src/b2_body.rs:91-147
FIXTURE = {"m_friction", "m_restitution", "m_restitution_threshold", "m_density", "m_is_sensor", "m_filter", "m_shape_type", "m_shape"}
SHAPES = {"ECircle": {"base", "m_p"}, "EEdge": {"base", "m_vertex1", "m_vertex2", "m_vertex0", "m_vertex3", "m_one_sided"},
          "EPolygon": {"base", "m_centroid", "m_count", "m_vertices", "m_normals"},
          "EChain": {"base", "m_vertices", "m_prev_vertex", "m_next_vertex"}}
JOINTS = {  # src/joints/serialize/*.rs
    "ERevoluteJoint": {"base", "local_anchor_a", "local_anchor_b", "reference_angle", "enable_limit", "lower_angle", "upper_angle",
                       "enable_motor", "motor_speed", "max_motor_torque"},
    "EPrismaticJoint": {"base", "local_anchor_a", "local_anchor_b", "local_axis_a", "reference_angle", "enable_limit",
                        "lower_translation", "upper_translation", "enable_motor", "motor_speed", "max_motor_force"},
    "EDistanceJoint": {"base", "local_anchor_a", "local_anchor_b", "length", "min_length", "max_length", "stiffness", "damping"},
    "EWeldJoint": {"base", "local_anchor_a", "local_anchor_b", "reference_angle", "stiffness", "damping"},
    "EWheelJoint": {"base", "local_anchor_a", "local_anchor_b", "local_axis_a", "enable_limit", "lower_translation",
                    "upper_translation", "enable_motor", "motor_speed", "max_motor_torque", "stiffness", "damping"},
    "EFrictionJoint": {"base", "local_anchor_a", "local_anchor_b", "max_force", "max_torque"},
    "EMotorJoint": {"base", "linear_offset", "angular_offset", "max_force", "max_torque", "correction_factor"},
    "EPulleyJoint": {"base", "ground_anchor_a", "ground_anchor_b", "local_anchor_a", "local_anchor_b", "length_a", "length_b", "ratio"},
    "EGearJoint": {"base", "joint1", "joint2", "ratio"},
}
ALL = dict(SCENES)
ALL.update(JOINT_SCENES)
NAMES = ["hello_world", "pyramid", "variety", "sensors", "terrain", "bridge", "joints_mix", "cantilever", "sliders", "car", "top_down", "pulleys", "gears"]


def _built(name):
    from box2d_rs_b200 import scenes
    from oracle import b2o
    recipe, gravity, _ = ALL[name]
    w = b2o.B2world(gravity)
    recipe(scenes, w)
    return w


@pytest.mark.parametrize("name", NAMES)
def test_document_layout(name, built):
    from box2d_rs_b200 import serde_world
    snap = _built(name).snapshot()
    doc = json.loads(json.dumps(serde_world.to_serde(snap)))
    assert set(doc) == {"m_gravity", "m_bodies_list", "m_joints_list", "m_gear_joints_list"}
    assert len(doc["m_bodies_list"]) == len(snap.bodies)
    for b in doc["m_bodies_list"]:
        assert set(b) == {"m_definition", "m_fixture_list"} and set(b["m_definition"]) == BODY_DEF
        for f in b["m_fixture_list"]:
            assert set(f) == FIXTURE and set(f["m_shape"]) == SHAPES[f["m_shape_type"]]
            assert set(f["m_shape"]["base"]) == {"m_type", "m_radius"} and f["m_shape"]["base"]["m_type"] == f["m_shape_type"]
            if f["m_shape_type"] == "EPolygon":
                assert len(f["m_shape"]["m_vertices"]) == len(f["m_shape"]["m_normals"]) == f["m_shape"]["m_count"]
    n_mouse = int((snap.joints["type"] == 5).sum()) if len(snap.joints) else 0
    assert len(doc["m_joints_list"]) + len(doc["m_gear_joints_list"]) == len(snap.joints) - n_mouse  # the crate skips mouse joints
    for e in doc["m_joints_list"] + doc["m_gear_joints_list"]:
        assert set(e) == {"jtype", "joint_def"} and set(e["joint_def"]) == JOINTS[e["jtype"]]
        base = e["joint_def"]["base"]
        assert set(base) == {"jtype", "user_data", "body_a", "body_b", "collide_connected"} and base["jtype"] == e["jtype"]
        assert 0 <= base["body_a"] < len(snap.bodies) and 0 <= base["body_b"] < len(snap.bodies)
    assert all(e["jtype"] == "EGearJoint" for e in doc["m_gear_joints_list"])
    # newest body first: the first document body is the last one created
    last = snap.bodies[len(snap.bodies) - 1]
    assert doc["m_bodies_list"][0]["m_definition"]["position"] == {"x": float(last["xf"][0]), "y": float(last["xf"][1])}


@pytest.mark.parametrize("name", [n for n in NAMES if n not in ("pulleys", "gears")])
def test_two_loads_restore_the_world(name, built):
    """One load reverses the creation order (bodies, fixtures, joints: every list is a push_front list), a second one restores
    it: the twice-loaded world equals the as-built one in every table — bodies, mass data, fixtures, shapes, proxies, tree."""
    from box2d_rs_b200 import serde_world
    from oracle import b2o
    w0 = _built(name)
    g = ALL[name][1]
    w1 = b2o.B2world((9.0, 9.0))
    serde_world.from_serde(json.loads(json.dumps(serde_world.to_serde(w0.snapshot()))), w1)
    w2 = b2o.B2world((9.0, 9.0))
    serde_world.from_serde(serde_world.to_serde(w1.snapshot()), w2)
    assert parity.compare_snapshots(w0.snapshot(), w2.snapshot()) == []
    assert w1.get_body_count() == w0.get_body_count() and w1.get_joint_count() == w0.get_joint_count()
    # the once-loaded world is the same scene listed backwards: same definitions, reversed
    d0, d1 = serde_world.to_serde(w0.snapshot()), serde_world.to_serde(w1.snapshot())
    assert [b["m_definition"] for b in d1["m_bodies_list"]] == [b["m_definition"] for b in d0["m_bodies_list"]][::-1]
    assert d1["m_gravity"] == d0["m_gravity"] == {"x": float(np.float32(g[0])), "y": float(np.float32(g[1]))}


def test_gear_and_pulley_scenes_reload(built):
    """Gear joints are written after the other joints (m_gear_joints_list) and mouse joints not at all, so these scenes come
    back with another joint order / without the drag: compare what a definition file can carry."""
    from box2d_rs_b200 import abi, scenes, serde_world
    from oracle import b2o
    for name in ("gears", "pulleys"):
        w0 = _built(name)
        w1 = b2o.B2world((0.0, 0.0))
        serde_world.from_serde(serde_world.to_serde(w0.snapshot()), w1)
        w2 = b2o.B2world((0.0, 0.0))
        serde_world.from_serde(serde_world.to_serde(w1.snapshot()), w2)
        s0, s2 = w0.snapshot(), w2.snapshot()
        assert np.array_equal(s0.bodies, s2.bodies) and np.array_equal(s0.fixtures, s2.fixtures) and np.array_equal(s0.shapes, s2.shapes)
        keep = [j for j in s0.joints if j["type"] != abi.JOINT_MOUSE]
        key = lambda j: j.tobytes()  # noqa: E731
        assert sorted(map(key, keep)) == sorted(map(key, s2.joints))
        if name == "gears":  # and the reloaded gear train still turns as the original does (same joints, another list order)
            for w in (w0, w2):
                for _ in range(30):
                    w.step(scenes.DT, 8, 3)
            a0 = np.sort(w0.snapshot().bodies["a"])
            a2 = np.sort(w2.snapshot().bodies["a"])
            assert np.allclose(a0, a2, atol=2e-3)


def test_both_mirrors_build_the_same_world(built):
    from box2d_rs_b200 import batch, scenes, serde_world, world
    from oracle import b2o
    ctx = batch.Context(0, lib_path=HOSTSIM_SO)
    for name in ("variety", "joints_mix", "gears", "car"):
        doc = json.loads(json.dumps(serde_world.to_serde(_built(name).snapshot())))
        wo = b2o.B2world((0.0, 0.0))
        serde_world.from_serde(doc, wo)
        wg = world.B2world((0.0, 0.0), ctx=ctx)
        serde_world.from_serde(doc, wg)
        assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
        assert serde_world.to_serde(wg.snapshot()) == serde_world.to_serde(wo.snapshot())
        for _ in range(40):
            wo.step(scenes.DT, 8, 3)
            wg.step(scenes.DT, 8, 3)
        assert parity.compare_snapshots(wo.snapshot(), wg.snapshot()) == []
        wg.close()
    ctx.close()
