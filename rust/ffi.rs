//! src/private/gpu/ffi.rs — the crate's only `unsafe`: the `extern "C"` surface of libb2gpu.so
//! (include/b2gpu.h, ABI version 2) and thin safe wrappers that turn error codes into `Result`.
//! NOT BUILT IN THIS REPOSITORY'S CI (no Rust toolchain in the image); kept in sync with the header by
//! tests/test_abi.py::test_rust_ffi_lists_every_symbol.
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_float, c_int, c_void};

pub const B2GPU_ABI_VERSION: c_int = 2;
pub const B2GPU_E_INVALID: c_int = -1;
pub const B2GPU_E_NO_DEVICE: c_int = -2;
pub const B2GPU_E_CUDA: c_int = -3;
pub const B2GPU_E_CAPACITY: c_int = -4;
pub const B2GPU_E_UNSUPPORTED: c_int = -5;
pub const B2GPU_E_LOCKED: c_int = -6;

#[repr(C)] pub struct b2gpu_ctx { _private: [u8; 0] }
#[repr(C)] pub struct b2gpu_world { _private: [u8; 0] }
#[repr(C)] pub struct b2gpu_batch { _private: [u8; 0] }

/// B2body fields the step reads or writes (128 bytes).
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_body_rec {
    pub type_: i32, pub flags: u32,
    pub xf_px: f32, pub xf_py: f32, pub xf_qs: f32, pub xf_qc: f32,
    pub lc_x: f32, pub lc_y: f32, pub c0_x: f32, pub c0_y: f32, pub c_x: f32, pub c_y: f32, pub a0: f32, pub a: f32,
    pub vx: f32, pub vy: f32, pub w: f32, pub fx: f32, pub fy: f32, pub torque: f32,
    pub mass: f32, pub inv_mass: f32, pub inertia: f32, pub inv_inertia: f32,
    pub linear_damping: f32, pub angular_damping: f32, pub gravity_scale: f32, pub sleep_time: f32,
    pub fixture_head: i32, pub fixture_count: i32, pub reserved: [i32; 2],
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_fixture_rec {
    pub body: i32, pub next: i32, pub shape_type: i32, pub shape_first: i32, pub child_count: i32, pub proxy_first: i32,
    pub density: f32, pub friction: f32, pub restitution: f32, pub restitution_threshold: f32,
    pub category_bits: u16, pub mask_bits: u16, pub group_index: i16, pub is_sensor: u16,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct b2gpu_shape_rec {
    pub type_: i32, pub radius: f32, pub count: i32, pub one_sided: i32, pub cx: f32, pub cy: f32,
    pub v: [f32; 16], pub n: [f32; 16], pub reserved: [i32; 2],
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_proxy_rec { pub fixture: i32, pub child_index: i32, pub proxy_id: i32, pub reserved: i32, pub aabb: [f32; 4] }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_tree_node_rec {
    pub aabb: [f32; 4], pub parent: i32, pub child1: i32, pub child2: i32, pub height: i32, pub proxy: i32, pub moved: i32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_manifold_point { pub lp_x: f32, pub lp_y: f32, pub normal_impulse: f32, pub tangent_impulse: f32, pub id: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_manifold {
    pub points: [b2gpu_manifold_point; 2], pub ln_x: f32, pub ln_y: f32, pub lp_x: f32, pub lp_y: f32,
    pub type_: i32, pub point_count: i32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_contact_rec {
    pub fixture_a: i32, pub fixture_b: i32, pub index_a: i32, pub index_b: i32, pub flags: u32,
    pub friction: f32, pub restitution: f32, pub restitution_threshold: f32, pub tangent_speed: f32, pub reserved: i32,
    pub manifold: b2gpu_manifold,
}
/// B2joint + B2revoluteJoint / B2distanceJoint: definition, parameters and accumulated impulses (96 bytes).
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_joint_rec {
    pub type_: i32, pub body_a: i32, pub body_b: i32, pub flags: u32,
    pub local_anchor_a: [f32; 2], pub local_anchor_b: [f32; 2], pub param: [f32; 8], pub impulse: [f32; 8],
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_joint_def {
    pub type_: i32, pub body_a: i32, pub body_b: i32, pub collide_connected: i32,
    pub local_anchor_a: [f32; 2], pub local_anchor_b: [f32; 2],
    pub reference_angle: f32, pub lower_angle: f32, pub upper_angle: f32, pub max_motor_torque: f32, pub motor_speed: f32,
    pub enable_limit: i32, pub enable_motor: i32,
    pub length: f32, pub min_length: f32, pub max_length: f32, pub stiffness: f32, pub damping: f32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_world_rec {
    pub gravity_x: f32, pub gravity_y: f32, pub inv_dt0: f32, pub flags: u32,
    pub tree_root: i32, pub tree_free_list: i32, pub tree_node_count: i32, pub tree_node_capacity: i32,
    pub tree_insertion_count: i32, pub proxy_count: i32, pub reserved: [i32; 2],
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_snapshot_sizes {
    pub body_count: i32, pub fixture_count: i32, pub shape_count: i32, pub proxy_count: i32,
    pub node_count: i32, pub contact_count: i32, pub move_count: i32, pub joint_count: i32,
}
#[repr(C)]
pub struct b2gpu_snapshot {
    pub world: b2gpu_world_rec, pub n: b2gpu_snapshot_sizes,
    pub bodies: *mut b2gpu_body_rec, pub fixtures: *mut b2gpu_fixture_rec, pub shapes: *mut b2gpu_shape_rec,
    pub proxies: *mut b2gpu_proxy_rec, pub nodes: *mut b2gpu_tree_node_rec, pub contacts: *mut b2gpu_contact_rec,
    pub move_buffer: *mut i32,
    pub joints: *mut b2gpu_joint_rec,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_step_stats {
    pub status: i32, pub contacts: i32, pub touching: i32, pub destroyed: i32, pub islands: i32, pub island_bodies: i32,
    pub island_contacts: i32, pub moved: i32, pub pairs: i32, pub created: i32, pub awake_bodies: i32,
    pub solver_levels: i32, pub reserved: [i32; 4],
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_body_def {
    pub type_: i32, pub position_x: f32, pub position_y: f32, pub angle: f32,
    pub linear_velocity_x: f32, pub linear_velocity_y: f32, pub angular_velocity: f32,
    pub linear_damping: f32, pub angular_damping: f32,
    pub allow_sleep: i32, pub awake: i32, pub fixed_rotation: i32, pub bullet: i32, pub enabled: i32, pub gravity_scale: f32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_fixture_def {
    pub friction: f32, pub restitution: f32, pub restitution_threshold: f32, pub density: f32, pub is_sensor: i32,
    pub category_bits: u16, pub mask_bits: u16, pub group_index: i16, pub reserved: u16,
}
#[repr(C)]
pub struct b2gpu_shape_def {
    pub type_: i32, pub radius: f32, pub p_x: f32, pub p_y: f32,
    pub v0: [f32; 2], pub v1: [f32; 2], pub v2: [f32; 2], pub v3: [f32; 2], pub one_sided: i32,
    pub count: i32, pub centroid: [f32; 2], pub vertices: [f32; 16], pub normals: [f32; 16],
    pub chain_vertices: *const f32, pub chain_count: i32, pub chain_prev: [f32; 2], pub chain_next: [f32; 2],
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_mass_data { pub mass: f32, pub center_x: f32, pub center_y: f32, pub inertia: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct b2gpu_caps {
    pub max_bodies: i32, pub max_fixtures: i32, pub max_shapes: i32, pub max_proxies: i32, pub max_contacts: i32,
    pub max_pairs: i32, pub reserved: [i32; 2],
}

/// b2gpu_ray_hit (include/b2gpu.h): closest hit of one ray, fixture == -1 when nothing was hit.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b2gpu_ray_hit {
    pub fixture: i32,
    pub child_index: i32,
    pub fraction: c_float,
    pub point: [c_float; 2],
    pub normal: [c_float; 2],
    pub reserved: i32,
}

/// b2gpu_contact_event (include/b2gpu.h): one begin_contact (1) / end_contact (2) of a step, in firing order.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b2gpu_contact_event {
    pub type_: i32,
    pub fixture_a: i32,
    pub index_a: i32,
    pub fixture_b: i32,
    pub index_b: i32,
    pub reserved: [i32; 3],
}

/// b2gpu_post_solve_event (include/b2gpu.h): one B2contactListener::post_solve report of the last step.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b2gpu_post_solve_event {
    pub fixture_a: i32, pub index_a: i32, pub fixture_b: i32, pub index_b: i32, pub count: i32,
    pub normal_impulses: [c_float; 2], pub tangent_impulses: [c_float; 2], pub reserved: [i32; 3],
}

extern "C" {
    pub fn b2gpu_abi_version() -> c_int;
    pub fn b2gpu_last_error() -> *const c_char;
    pub fn b2gpu_device_count() -> c_int;
    pub fn b2gpu_init(device: c_int, stream: *mut c_void, out: *mut *mut b2gpu_ctx) -> c_int;
    pub fn b2gpu_shutdown(ctx: *mut b2gpu_ctx);
    pub fn b2gpu_sync(ctx: *mut b2gpu_ctx) -> c_int;
    pub fn b2gpu_stream(ctx: *mut b2gpu_ctx) -> *mut c_void;
    pub fn b2gpu_launch_count(ctx: *mut b2gpu_ctx) -> i64;
    pub fn b2gpu_polygon_set_as_box(s: *mut b2gpu_shape_def, hx: c_float, hy: c_float) -> c_int;
    pub fn b2gpu_polygon_set_as_box_angle(s: *mut b2gpu_shape_def, hx: c_float, hy: c_float, cx: c_float, cy: c_float, angle: c_float) -> c_int;
    pub fn b2gpu_polygon_set(s: *mut b2gpu_shape_def, vertices_xy: *const c_float, count: c_int) -> c_int;
    pub fn b2gpu_shape_compute_mass(s: *const b2gpu_shape_def, density: c_float, out: *mut b2gpu_mass_data) -> c_int;
    pub fn b2gpu_world_create(ctx: *mut b2gpu_ctx, gravity_x: c_float, gravity_y: c_float, out: *mut *mut b2gpu_world) -> c_int;
    pub fn b2gpu_world_destroy(w: *mut b2gpu_world);
    pub fn b2gpu_world_create_body(w: *mut b2gpu_world, def: *const b2gpu_body_def) -> c_int;
    pub fn b2gpu_body_create_fixture(w: *mut b2gpu_world, body: c_int, def: *const b2gpu_fixture_def, shape: *const b2gpu_shape_def) -> c_int;
    pub fn b2gpu_body_set_transform(w: *mut b2gpu_world, body: c_int, px: c_float, py: c_float, angle: c_float) -> c_int;
    pub fn b2gpu_body_set_linear_velocity(w: *mut b2gpu_world, body: c_int, vx: c_float, vy: c_float) -> c_int;
    pub fn b2gpu_body_set_angular_velocity(w: *mut b2gpu_world, body: c_int, w_: c_float) -> c_int;
    pub fn b2gpu_body_apply_force_to_center(w: *mut b2gpu_world, body: c_int, fx: c_float, fy: c_float, wake: c_int) -> c_int;
    pub fn b2gpu_body_apply_force(w: *mut b2gpu_world, body: c_int, fx: c_float, fy: c_float, point_x: c_float, point_y: c_float, wake: c_int) -> c_int;
    pub fn b2gpu_body_apply_torque(w: *mut b2gpu_world, body: c_int, torque: c_float, wake: c_int) -> c_int;
    pub fn b2gpu_body_apply_linear_impulse(w: *mut b2gpu_world, body: c_int, ix: c_float, iy: c_float, point_x: c_float, point_y: c_float, wake: c_int) -> c_int;
    pub fn b2gpu_body_apply_linear_impulse_to_center(w: *mut b2gpu_world, body: c_int, ix: c_float, iy: c_float, wake: c_int) -> c_int;
    pub fn b2gpu_body_apply_angular_impulse(w: *mut b2gpu_world, body: c_int, impulse: c_float, wake: c_int) -> c_int;
    pub fn b2gpu_body_set_awake(w: *mut b2gpu_world, body: c_int, flag: c_int) -> c_int;
    pub fn b2gpu_body_set_damping(w: *mut b2gpu_world, body: c_int, linear_damping: c_float, angular_damping: c_float) -> c_int;
    pub fn b2gpu_body_set_gravity_scale(w: *mut b2gpu_world, body: c_int, scale: c_float) -> c_int;
    pub fn b2gpu_body_set_sleeping_allowed(w: *mut b2gpu_world, body: c_int, flag: c_int) -> c_int;
    pub fn b2gpu_revolute_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int, anchor_x: c_float, anchor_y: c_float) -> c_int;
    pub fn b2gpu_distance_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int, a1x: c_float, a1y: c_float, a2x: c_float, a2y: c_float) -> c_int;
    pub fn b2gpu_prismatic_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int, anchor_x: c_float, anchor_y: c_float, axis_x: c_float, axis_y: c_float) -> c_int;
    pub fn b2gpu_friction_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int, anchor_x: c_float, anchor_y: c_float) -> c_int;
    pub fn b2gpu_motor_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int) -> c_int;
    pub fn b2gpu_pulley_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int,
                                  ground_ax: c_float, ground_ay: c_float, ground_bx: c_float, ground_by: c_float,
                                  anchor_ax: c_float, anchor_ay: c_float, anchor_bx: c_float, anchor_by: c_float,
                                  ratio: c_float) -> c_int;
    pub fn b2gpu_gear_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, joint1: c_int, joint2: c_int, ratio: c_float) -> c_int;
    pub fn b2gpu_mouse_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int,
                                 target_x: c_float, target_y: c_float) -> c_int;
    pub fn b2gpu_wheel_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int, anchor_x: c_float, anchor_y: c_float, axis_x: c_float, axis_y: c_float) -> c_int;
    pub fn b2gpu_weld_joint_def(w: *mut b2gpu_world, def: *mut b2gpu_joint_def, body_a: c_int, body_b: c_int, anchor_x: c_float, anchor_y: c_float) -> c_int;
    pub fn b2gpu_angular_stiffness(w: *mut b2gpu_world, frequency_hertz: c_float, damping_ratio: c_float, body_a: c_int, body_b: c_int, stiffness: *mut c_float, damping: *mut c_float) -> c_int;
    pub fn b2gpu_linear_stiffness(w: *mut b2gpu_world, frequency_hertz: c_float, damping_ratio: c_float, body_a: c_int, body_b: c_int, stiffness: *mut c_float, damping: *mut c_float) -> c_int;
    pub fn b2gpu_world_create_joint(w: *mut b2gpu_world, def: *const b2gpu_joint_def) -> c_int;
    pub fn b2gpu_world_get_joint_count(w: *mut b2gpu_world) -> c_int;
    pub fn b2gpu_world_get_joint(w: *mut b2gpu_world, joint: c_int, out: *mut b2gpu_joint_rec) -> c_int;
    pub fn b2gpu_joint_set_motor_speed(w: *mut b2gpu_world, joint: c_int, speed: c_float) -> c_int;
    pub fn b2gpu_joint_set_max_motor_torque(w: *mut b2gpu_world, joint: c_int, torque: c_float) -> c_int;
    pub fn b2gpu_joint_enable_motor(w: *mut b2gpu_world, joint: c_int, flag: c_int) -> c_int;
    pub fn b2gpu_joint_enable_limit(w: *mut b2gpu_world, joint: c_int, flag: c_int) -> c_int;
    pub fn b2gpu_joint_set_limits(w: *mut b2gpu_world, joint: c_int, lower: c_float, upper: c_float) -> c_int;
    pub fn b2gpu_world_set_gravity(w: *mut b2gpu_world, gravity_x: c_float, gravity_y: c_float) -> c_int;
    pub fn b2gpu_world_get_gravity(w: *mut b2gpu_world, gravity_x: *mut c_float, gravity_y: *mut c_float) -> c_int;
    pub fn b2gpu_world_destroy_joint(w: *mut b2gpu_world, joint: c_int) -> c_int;
    pub fn b2gpu_joint_set_target(w: *mut b2gpu_world, joint: c_int, target_x: c_float, target_y: c_float) -> c_int;
    pub fn b2gpu_world_set_allow_sleeping(w: *mut b2gpu_world, flag: c_int) -> c_int;
    pub fn b2gpu_world_set_warm_starting(w: *mut b2gpu_world, flag: c_int) -> c_int;
    pub fn b2gpu_world_set_continuous_physics(w: *mut b2gpu_world, flag: c_int) -> c_int;
    pub fn b2gpu_world_set_block_solve(w: *mut b2gpu_world, flag: c_int) -> c_int;
    pub fn b2gpu_world_set_large_mode(w: *mut b2gpu_world, flag: c_int) -> c_int;
    pub fn b2gpu_world_set_level_threshold(w: *mut b2gpu_world, contacts: c_int) -> c_int;
    pub fn b2gpu_world_ray_cast_closest(w: *mut b2gpu_world, p1p2: *const c_float, n: c_int, out: *mut b2gpu_ray_hit) -> c_int;
    pub fn b2gpu_world_query_aabb(w: *mut b2gpu_world, aabbs: *const c_float, n: c_int, max_hits: c_int, counts: *mut i32, hits: *mut i32) -> c_int;
    pub fn b2gpu_batch_ray_cast_closest(b: *mut b2gpu_batch, p1p2: *const c_float, rays_per_world: c_int, out: *mut b2gpu_ray_hit) -> c_int;
    pub fn b2gpu_batch_query_aabb(b: *mut b2gpu_batch, aabbs: *const f32, boxes_per_world: c_int, max_hits: c_int, counts: *mut i32, hits: *mut i32) -> c_int;
    pub fn b2gpu_world_step(w: *mut b2gpu_world, dt: c_float, velocity_iterations: c_int, position_iterations: c_int) -> c_int;
    pub fn b2gpu_world_get_body_count(w: *mut b2gpu_world) -> c_int;
    pub fn b2gpu_world_get_contact_count(w: *mut b2gpu_world) -> c_int;
    pub fn b2gpu_world_get_body(w: *mut b2gpu_world, body: c_int, out: *mut b2gpu_body_rec) -> c_int;
    pub fn b2gpu_world_get_stats(w: *mut b2gpu_world, out: *mut b2gpu_step_stats) -> c_int;
    pub fn b2gpu_world_snapshot_sizes(w: *mut b2gpu_world, out: *mut b2gpu_snapshot_sizes) -> c_int;
    pub fn b2gpu_world_download(w: *mut b2gpu_world, out: *mut b2gpu_snapshot) -> c_int;
    pub fn b2gpu_world_upload(w: *mut b2gpu_world, snap: *const b2gpu_snapshot) -> c_int;
    // checkpoint / resume (host-only): the whole step state on disk, where the crate's serde support
    // (src/serialize/serialize_b2_world.rs) saves definitions only
    pub fn b2gpu_contact_events(before: *const b2gpu_snapshot, after: *const b2gpu_snapshot, destroyed: c_int, out: *mut b2gpu_contact_event, capacity: c_int) -> c_int;
    pub fn b2gpu_snapshot_validate(s: *const b2gpu_snapshot) -> c_int;
    pub fn b2gpu_snapshot_save(s: *const b2gpu_snapshot, path: *const c_char) -> c_int;
    pub fn b2gpu_snapshot_file_sizes(path: *const c_char, out: *mut b2gpu_snapshot_sizes) -> c_int;
    pub fn b2gpu_snapshot_load(path: *const c_char, out: *mut b2gpu_snapshot) -> c_int;
    pub fn b2gpu_batch_create(ctx: *mut b2gpu_ctx, proto: *const b2gpu_snapshot, n_worlds: c_int, caps: *const b2gpu_caps, out: *mut *mut b2gpu_batch) -> c_int;
    pub fn b2gpu_batch_destroy(b: *mut b2gpu_batch);
    pub fn b2gpu_batch_world_count(b: *mut b2gpu_batch) -> c_int;
    pub fn b2gpu_batch_step(b: *mut b2gpu_batch, dt: c_float, velocity_iterations: c_int, position_iterations: c_int, steps: c_int) -> c_int;
    pub fn b2gpu_batch_upload_world(b: *mut b2gpu_batch, world: c_int, snap: *const b2gpu_snapshot) -> c_int;
    pub fn b2gpu_batch_snapshot_sizes(b: *mut b2gpu_batch, world: c_int, out: *mut b2gpu_snapshot_sizes) -> c_int;
    pub fn b2gpu_batch_download_world(b: *mut b2gpu_batch, world: c_int, out: *mut b2gpu_snapshot) -> c_int;
    pub fn b2gpu_batch_get_stats(b: *mut b2gpu_batch, first_world: c_int, count: c_int, out: *mut b2gpu_step_stats) -> c_int;
    pub fn b2gpu_world_post_solve_events(w: *mut b2gpu_world, out: *mut b2gpu_post_solve_event, capacity: c_int) -> c_int;
    pub fn b2gpu_batch_post_solve_events(b: *mut b2gpu_batch, world: c_int, out: *mut b2gpu_post_solve_event, capacity: c_int) -> c_int;
    pub fn b2gpu_batch_reset(b: *mut b2gpu_batch, input: *const b2gpu_snapshot) -> c_int;
    pub fn b2gpu_batch_status(b: *mut b2gpu_batch) -> c_int;
    pub fn b2gpu_batch_set_level_threshold(b: *mut b2gpu_batch, contacts: c_int) -> c_int;
    pub fn b2gpu_batch_set_forces(b: *mut b2gpu_batch, host_fxfyt: *const c_float, first_world: c_int, count: c_int) -> c_int;
    pub fn b2gpu_batch_set_gravity(b: *mut b2gpu_batch, host_gxgy: *const c_float, first_world: c_int, count: c_int) -> c_int;
    pub fn b2gpu_batch_set_joint_control(b: *mut b2gpu_batch, joint: c_int, control: c_int, host_values: *const c_float, first_world: c_int, count: c_int) -> c_int;
    pub fn b2gpu_batch_set_linear_velocity(b: *mut b2gpu_batch, body: c_int, host_vxvy: *const c_float, first_world: c_int, count: c_int) -> c_int;
    pub fn b2gpu_batch_get_body_state(b: *mut b2gpu_batch, host_out: *mut c_float, first_world: c_int, count: c_int) -> c_int;
    pub fn b2gpu_batch_body_state_device(b: *mut b2gpu_batch, bytes: *mut i64) -> *mut c_void;
    pub fn b2gpu_batch_forces_device(b: *mut b2gpu_batch, bytes: *mut i64) -> *mut c_void;
    pub fn b2gpu_batch_apply_device_forces(b: *mut b2gpu_batch) -> c_int;
    pub fn b2gpu_batch_refresh_device_state(b: *mut b2gpu_batch) -> c_int;
    pub fn b2gpu_batch_step_host(b: *mut b2gpu_batch, host_forces: *const c_float, host_state_out: *mut c_float, dt: c_float,
                                 velocity_iterations: c_int, position_iterations: c_int, steps: c_int) -> c_int;
    pub fn b2gpu_batch_dynamic_bodies(b: *mut b2gpu_batch, out_body_indices: *mut i32, capacity: c_int) -> c_int;
    pub fn b2gpu_batch_step_host_dynamic(b: *mut b2gpu_batch, host_forces: *const c_float, host_state_out: *mut c_float, dt: c_float, velocity_iterations: c_int, position_iterations: c_int, steps: c_int) -> c_int;
    pub fn b2gpu_batch_algorithmic_bytes(b: *mut b2gpu_batch) -> i64;
    pub fn b2gpu_stage_count() -> c_int;
    pub fn b2gpu_stage_name(stage: c_int) -> *const c_char;
    pub fn b2gpu_set_profiling(ctx: *mut b2gpu_ctx, on: c_int) -> c_int;
    pub fn b2gpu_get_stage_times(ctx: *mut b2gpu_ctx, ms_out: *mut f64, launches_out: *mut i64, n: c_int) -> c_int;
    pub fn b2gpu_debug_sincos(ctx: *mut b2gpu_ctx, host_in: *const c_float, host_sin: *mut c_float, host_cos: *mut c_float, n: c_int) -> c_int;
}

fn check(rc: c_int) -> Result<c_int, String> {
    if rc >= 0 { return Ok(rc); }
    // SAFETY: b2gpu_last_error returns a NUL-terminated string owned by the library (thread-local).
    let msg = unsafe { CStr::from_ptr(b2gpu_last_error()) }.to_string_lossy().into_owned();
    Err(format!("b2gpu error {rc}: {msg}"))
}

/// One context per device.  `!Send` like the reference world (Rc<RefCell<..>>).
pub struct Context { raw: *mut b2gpu_ctx }
impl Context {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        // SAFETY: `raw` is a valid out-pointer; on failure the library leaves it null.
        check(unsafe { b2gpu_init(device, std::ptr::null_mut(), &mut raw) })?;
        Ok(Context { raw })
    }
}
impl Drop for Context {
    fn drop(&mut self) { unsafe { b2gpu_shutdown(self.raw) } }
}

/// The GPU mirror of one B2world; `step` replaces the body of private::step (b2_world.rs(private):903-959).
pub struct GpuWorld<'c> { raw: *mut b2gpu_world, _ctx: &'c Context }
impl<'c> GpuWorld<'c> {
    pub fn new(ctx: &'c Context, gravity: (f32, f32)) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { b2gpu_world_create(ctx.raw, gravity.0, gravity.1, &mut raw) })?;
        Ok(GpuWorld { raw, _ctx: ctx })
    }
    pub fn upload(&mut self, snap: &b2gpu_snapshot) -> Result<(), String> {
        check(unsafe { b2gpu_world_upload(self.raw, snap) }).map(|_| ())
    }
    pub fn step(&mut self, dt: f32, velocity_iterations: i32, position_iterations: i32) -> Result<(), String> {
        check(unsafe { b2gpu_world_step(self.raw, dt, velocity_iterations, position_iterations) }).map(|_| ())
    }
    pub fn snapshot_sizes(&mut self) -> Result<b2gpu_snapshot_sizes, String> {
        let mut n = b2gpu_snapshot_sizes::default();
        check(unsafe { b2gpu_world_snapshot_sizes(self.raw, &mut n) })?;
        Ok(n)
    }
    /// Caller provides buffers sized by `snapshot_sizes` (capacities in `out.n`).
    pub fn download(&mut self, out: &mut b2gpu_snapshot) -> Result<(), String> {
        check(unsafe { b2gpu_world_download(self.raw, out) }).map(|_| ())
    }
}
impl Drop for GpuWorld<'_> {
    fn drop(&mut self) { unsafe { b2gpu_world_destroy(self.raw) } }
}
