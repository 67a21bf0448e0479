"""ctypes binding of the C ABI declared in include/b2gpu.h (libb2gpu.so, built in-tree by
__graft_entry__.build()).  There is no fallback: a missing library raises at load time, and on a
machine without a CUDA device every stepping call returns B2GPU_E_NO_DEVICE."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.path.join(_HERE, "libb2gpu.so")
_LIBS = {}


class B2gpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b2gpu error %d: %s" % (code, msg))
        self.code = code


def load(path=None):
    path = os.path.abspath(path or DEFAULT_SO)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()')" % path)
    L = C.CDLL(path)
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    sig = {
        "b2gpu_abi_version": (i32, []),
        "b2gpu_last_error": (C.c_char_p, []),
        "b2gpu_device_count": (i32, []),
        "b2gpu_init": (i32, [i32, vp, C.POINTER(vp)]),
        "b2gpu_shutdown": (None, [vp]),
        "b2gpu_sync": (i32, [vp]),
        "b2gpu_stream": (vp, [vp]),
        "b2gpu_launch_count": (i64, [vp]),
        "b2gpu_polygon_set_as_box": (i32, [C.POINTER(abi.ShapeDef), f32, f32]),
        "b2gpu_polygon_set_as_box_angle": (i32, [C.POINTER(abi.ShapeDef), f32, f32, f32, f32, f32]),
        "b2gpu_polygon_set": (i32, [C.POINTER(abi.ShapeDef), C.POINTER(f32), i32]),
        "b2gpu_shape_compute_mass": (i32, [C.POINTER(abi.ShapeDef), f32, C.POINTER(abi.MassData)]),
        "b2gpu_world_create": (i32, [vp, f32, f32, C.POINTER(vp)]),
        "b2gpu_world_destroy": (None, [vp]),
        "b2gpu_world_create_body": (i32, [vp, C.POINTER(abi.BodyDef)]),
        "b2gpu_body_create_fixture": (i32, [vp, i32, C.POINTER(abi.FixtureDef), C.POINTER(abi.ShapeDef)]),
        "b2gpu_body_set_transform": (i32, [vp, i32, f32, f32, f32]),
        "b2gpu_body_set_linear_velocity": (i32, [vp, i32, f32, f32]),
        "b2gpu_body_set_angular_velocity": (i32, [vp, i32, f32]),
        "b2gpu_body_apply_force_to_center": (i32, [vp, i32, f32, f32, i32]),
        "b2gpu_body_apply_force": (i32, [vp, i32, f32, f32, f32, f32, i32]),
        "b2gpu_body_apply_torque": (i32, [vp, i32, f32, i32]),
        "b2gpu_body_apply_linear_impulse": (i32, [vp, i32, f32, f32, f32, f32, i32]),
        "b2gpu_body_apply_linear_impulse_to_center": (i32, [vp, i32, f32, f32, i32]),
        "b2gpu_body_apply_angular_impulse": (i32, [vp, i32, f32, i32]),
        "b2gpu_body_set_awake": (i32, [vp, i32, i32]),
        "b2gpu_body_set_damping": (i32, [vp, i32, f32, f32]),
        "b2gpu_body_set_gravity_scale": (i32, [vp, i32, f32]),
        "b2gpu_body_set_sleeping_allowed": (i32, [vp, i32, i32]),
        "b2gpu_revolute_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32]),
        "b2gpu_distance_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32, f32, f32]),
        "b2gpu_linear_stiffness": (i32, [vp, f32, f32, i32, i32, C.POINTER(f32), C.POINTER(f32)]),
        "b2gpu_weld_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32]),
        "b2gpu_prismatic_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32, f32, f32]),
        "b2gpu_wheel_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32, f32, f32]),
        "b2gpu_friction_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32]),
        "b2gpu_motor_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32]),
        "b2gpu_pulley_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32] + [f32] * 9),
        "b2gpu_gear_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32]),
        "b2gpu_mouse_joint_def": (i32, [vp, C.POINTER(abi.JointDef), i32, i32, f32, f32]),
        "b2gpu_joint_set_target": (i32, [vp, i32, f32, f32]),
        "b2gpu_world_destroy_joint": (i32, [vp, i32]),
        "b2gpu_world_set_gravity": (i32, [vp, f32, f32]),
        "b2gpu_world_get_gravity": (i32, [vp, C.POINTER(f32), C.POINTER(f32)]),
        "b2gpu_angular_stiffness": (i32, [vp, f32, f32, i32, i32, C.POINTER(f32), C.POINTER(f32)]),
        "b2gpu_world_create_joint": (i32, [vp, C.POINTER(abi.JointDef)]),
        "b2gpu_world_get_joint_count": (i32, [vp]),
        "b2gpu_world_get_joint": (i32, [vp, i32, vp]),
        "b2gpu_joint_set_motor_speed": (i32, [vp, i32, f32]),
        "b2gpu_joint_set_max_motor_torque": (i32, [vp, i32, f32]),
        "b2gpu_joint_enable_motor": (i32, [vp, i32, i32]),
        "b2gpu_joint_enable_limit": (i32, [vp, i32, i32]),
        "b2gpu_joint_set_limits": (i32, [vp, i32, f32, f32]),
        "b2gpu_world_set_allow_sleeping": (i32, [vp, i32]),
        "b2gpu_world_set_warm_starting": (i32, [vp, i32]),
        "b2gpu_world_set_continuous_physics": (i32, [vp, i32]),
        "b2gpu_world_set_block_solve": (i32, [vp, i32]),
        "b2gpu_world_set_large_mode": (i32, [vp, i32]),
        "b2gpu_world_set_level_threshold": (i32, [vp, i32]),
        "b2gpu_world_step": (i32, [vp, f32, i32, i32]),
        "b2gpu_world_get_body_count": (i32, [vp]),
        "b2gpu_world_get_contact_count": (i32, [vp]),
        "b2gpu_world_get_body": (i32, [vp, i32, vp]),
        "b2gpu_world_get_stats": (i32, [vp, vp]),
        "b2gpu_world_snapshot_sizes": (i32, [vp, C.POINTER(abi.SnapshotSizes)]),
        "b2gpu_world_download": (i32, [vp, C.POINTER(abi.SnapshotC)]),
        "b2gpu_world_upload": (i32, [vp, C.POINTER(abi.SnapshotC)]),
        "b2gpu_snapshot_validate": (i32, [C.POINTER(abi.SnapshotC)]),
        "b2gpu_snapshot_save": (i32, [C.POINTER(abi.SnapshotC), C.c_char_p]),
        "b2gpu_snapshot_file_sizes": (i32, [C.c_char_p, C.POINTER(abi.SnapshotSizes)]),
        "b2gpu_snapshot_load": (i32, [C.c_char_p, C.POINTER(abi.SnapshotC)]),
        "b2gpu_world_post_solve_events": (i32, [vp, vp, i32]),
        "b2gpu_batch_post_solve_events": (i32, [vp, i32, vp, i32]),
        "b2gpu_contact_events": (i32, [C.POINTER(abi.SnapshotC), C.POINTER(abi.SnapshotC), i32, vp, i32]),
        "b2gpu_world_ray_cast_closest": (i32, [vp, vp, i32, vp]),
        "b2gpu_world_query_aabb": (i32, [vp, vp, i32, i32, vp, vp]),
        "b2gpu_batch_ray_cast_closest": (i32, [vp, vp, i32, vp]),
        "b2gpu_batch_query_aabb": (i32, [vp, vp, i32, i32, vp, vp]),
        "b2gpu_batch_create": (i32, [vp, C.POINTER(abi.SnapshotC), i32, C.POINTER(abi.Caps), C.POINTER(vp)]),
        "b2gpu_batch_destroy": (None, [vp]),
        "b2gpu_batch_world_count": (i32, [vp]),
        "b2gpu_batch_step": (i32, [vp, f32, i32, i32, i32]),
        "b2gpu_batch_upload_world": (i32, [vp, i32, C.POINTER(abi.SnapshotC)]),
        "b2gpu_batch_snapshot_sizes": (i32, [vp, i32, C.POINTER(abi.SnapshotSizes)]),
        "b2gpu_batch_download_world": (i32, [vp, i32, C.POINTER(abi.SnapshotC)]),
        "b2gpu_batch_get_stats": (i32, [vp, i32, i32, vp]),
        "b2gpu_batch_reset": (i32, [vp, C.POINTER(abi.SnapshotC)]),
        "b2gpu_batch_status": (i32, [vp]),
        "b2gpu_batch_set_level_threshold": (i32, [vp, i32]),
        "b2gpu_batch_set_forces": (i32, [vp, vp, i32, i32]),
        "b2gpu_batch_set_linear_velocity": (i32, [vp, i32, vp, i32, i32]),
        "b2gpu_batch_set_joint_control": (i32, [vp, i32, i32, vp, i32, i32]),
        "b2gpu_batch_set_gravity": (i32, [vp, vp, i32, i32]),
        "b2gpu_batch_get_body_state": (i32, [vp, vp, i32, i32]),
        "b2gpu_batch_body_state_device": (vp, [vp, C.POINTER(i64)]),
        "b2gpu_batch_forces_device": (vp, [vp, C.POINTER(i64)]),
        "b2gpu_batch_apply_device_forces": (i32, [vp]),
        "b2gpu_batch_refresh_device_state": (i32, [vp]),
        "b2gpu_batch_step_host": (i32, [vp, vp, vp, f32, i32, i32, i32]),
        "b2gpu_batch_dynamic_bodies": (i32, [vp, vp, i32]),
        "b2gpu_batch_step_host_dynamic": (i32, [vp, vp, vp, f32, i32, i32, i32]),
        "b2gpu_batch_algorithmic_bytes": (i64, [vp]),
        "b2gpu_debug_sincos": (i32, [vp, vp, vp, vp, i32]),
        "b2gpu_stage_count": (i32, []),
        "b2gpu_stage_name": (C.c_char_p, [i32]),
        "b2gpu_set_profiling": (i32, [vp, i32]),
        "b2gpu_get_stage_times": (i32, [vp, vp, vp, i32]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    L._b2gpu_missing = missing
    L._b2gpu_declared = sorted(sig)
    _LIBS[path] = L
    return L


def check(L, rc):
    if rc < 0:
        raise B2gpuError(rc, (L.b2gpu_last_error() or b"").decode("utf-8", "replace"))
    return rc
