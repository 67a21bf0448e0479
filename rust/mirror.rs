//! src/private/gpu/mirror.rs — the host mirror between box2d-rs's `Rc<RefCell<…>>` world graph and the flat snapshot
//! records of include/b2gpu.h.  Safe Rust only: the `unsafe` FFI lives in ffi.rs.
//!
//!   flatten(&world)            -> Snapshot      every table of b2gpu_snapshot in creation order
//!   write_back(&world, &snap)                   device results back into B2body / B2contact / tree / joints
//!   Snapshot::save(path)                        the checkpoint file of b2g_checkpoint.cu, written in pure Rust
//!
//! NOT COMPILED IN THIS REPOSITORY (the image has no Rust toolchain); it is written against the crate's fields as they
//! are at box2d-rs 0.0.4 and cites them.  It must live INSIDE the crate (`pub(crate)` fields): add
//! `pub mod gpu { pub mod ffi; pub mod mirror; }` to src/private/mod.rs and re-export `gpu::mirror::flatten` as
//! `B2world::gpu_snapshot`.
//!
//! Order contract: every intrusive list of the reference is a push_front list, so iterating a list and
//! reversing gives creation order.  Bodies: world list reversed.  Fixtures: body by body (creation order), each body's
//! list reversed — this equals global fixture-creation order whenever a body gets its fixtures before the next body is
//! created, which holds for every scene of this repository; indices are only names, the step itself does not depend on
//! the interleaving.  Proxies: fixture by fixture, children ascending.  Contacts / joints: world lists reversed.
#![allow(dead_code)]
use std::cell::RefCell;
use std::collections::HashMap;
use std::rc::Rc;

use super::ffi::*;
use crate::b2_body::*;
use crate::b2_collision::*;
use crate::b2_contact::*;
use crate::b2_fixture::*;
use crate::b2_joint::*;
use crate::b2_math::*;
use crate::b2_shape::*;
use crate::b2_world::*;
use crate::b2rs_common::UserDataType;
use crate::shapes::b2_edge_shape::B2edgeShape;
use crate::shapes::b2rs_to_derived_shape::ShapeAsDerived;

/// Owned tables of one snapshot plus the Rc handles they were flattened from (for write_back).
pub struct Snapshot<D: UserDataType> {
    pub world: b2gpu_world_rec,
    pub bodies: Vec<b2gpu_body_rec>,
    pub fixtures: Vec<b2gpu_fixture_rec>,
    pub shapes: Vec<b2gpu_shape_rec>,
    pub proxies: Vec<b2gpu_proxy_rec>,
    pub nodes: Vec<b2gpu_tree_node_rec>,
    pub contacts: Vec<b2gpu_contact_rec>,
    pub move_buffer: Vec<i32>,
    pub joints: Vec<b2gpu_joint_rec>,
    pub body_ptrs: Vec<BodyPtr<D>>,
    pub fixture_ptrs: Vec<FixturePtr<D>>,
    pub proxy_ptrs: Vec<FixtureProxyPtr<D>>,
    pub joint_ptrs: Vec<B2jointPtr<D>>,
}

fn addr<T: ?Sized>(rc: &Rc<RefCell<T>>) -> usize {
    Rc::as_ptr(rc) as *const () as usize
}

fn contact_id_key(id: &B2contactId) -> u32 {
    // include/b2gpu.h: index_a | index_b << 8 | type_a << 16 | type_b << 24 (src/b2_collision.rs:23-38)
    (id.cf.index_a as u32) | ((id.cf.index_b as u32) << 8) | ((id.cf.type_a as u32) << 16) | ((id.cf.type_b as u32) << 24)
}
fn contact_id_from_key(key: u32) -> B2contactId {
    B2contactId { cf: B2contactFeature { index_a: key as u8, index_b: (key >> 8) as u8, type_a: (key >> 16) as u8, type_b: (key >> 24) as u8 } }
}

fn edge_rec(e: &B2edgeShape) -> b2gpu_shape_rec {
    let mut r = b2gpu_shape_rec { type_: 1, radius: e.base.m_radius, count: 0, one_sided: e.m_one_sided as i32, cx: 0.0, cy: 0.0,
                                  v: [0.0; 16], n: [0.0; 16], reserved: [0; 2] };
    let vs = [e.m_vertex0, e.m_vertex1, e.m_vertex2, e.m_vertex3];
    for (i, p) in vs.iter().enumerate() { r.v[2 * i] = p.x; r.v[2 * i + 1] = p.y; }
    r
}

/// One record per collision child of a shape (a chain materialises its edges, b2_chain_shape.rs(private):59-79).
fn shape_recs(shape: &dyn B2shapeDynTrait, out: &mut Vec<b2gpu_shape_rec>) {
    match shape.as_derived() {
        ShapeAsDerived::AsCircle(c) => {
            let mut r = b2gpu_shape_rec { type_: 0, radius: c.base.m_radius, count: 0, one_sided: 0, cx: c.m_p.x, cy: c.m_p.y,
                                          v: [0.0; 16], n: [0.0; 16], reserved: [0; 2] };
            r.v[0] = c.m_p.x;
            r.v[1] = c.m_p.y;
            out.push(r);
        }
        ShapeAsDerived::AsEdge(e) => out.push(edge_rec(e)),
        ShapeAsDerived::AsPolygon(p) => {
            let mut r = b2gpu_shape_rec { type_: 2, radius: p.base.m_radius, count: p.m_count as i32, one_sided: 0,
                                          cx: p.m_centroid.x, cy: p.m_centroid.y, v: [0.0; 16], n: [0.0; 16], reserved: [0; 2] };
            for i in 0..8 {
                r.v[2 * i] = p.m_vertices[i].x; r.v[2 * i + 1] = p.m_vertices[i].y;
                r.n[2 * i] = p.m_normals[i].x; r.n[2 * i + 1] = p.m_normals[i].y;
            }
            out.push(r);
        }
        ShapeAsDerived::AsChain(c) => {
            for i in 0..shape.get_child_count() {
                let mut e = B2edgeShape::default();
                c.get_child_edge(&mut e, i);
                out.push(edge_rec(&e));
            }
        }
    }
}

/// The whole step state of `world` as flat tables (b2gpu_world_upload / b2gpu_batch_create input).
pub fn flatten<D: UserDataType>(world: &B2world<D>) -> Snapshot<D> {
    let cm = world.m_contact_manager.borrow();
    let bp = cm.m_broad_phase.borrow();
    let tree = &bp.m_tree;

    // ---- bodies (world list reversed = creation order)
    let mut body_ptrs: Vec<BodyPtr<D>> = world.m_body_list.iter().collect();
    body_ptrs.reverse();
    let body_index: HashMap<usize, i32> = body_ptrs.iter().enumerate().map(|(i, b)| (addr(b), i as i32)).collect();

    // ---- fixtures, shapes, proxies
    let mut fixture_ptrs: Vec<FixturePtr<D>> = Vec::new();
    let mut fixtures: Vec<b2gpu_fixture_rec> = Vec::new();
    let mut shapes: Vec<b2gpu_shape_rec> = Vec::new();
    let mut proxy_ptrs: Vec<FixtureProxyPtr<D>> = Vec::new();
    let mut proxies: Vec<b2gpu_proxy_rec> = Vec::new();
    let mut fixture_head: Vec<i32> = vec![-1; body_ptrs.len()];
    for (bi, b) in body_ptrs.iter().enumerate() {
        let mut own: Vec<FixturePtr<D>> = b.borrow().m_fixture_list.iter().collect();
        own.reverse();
        let mut prev: i32 = -1; // next-older fixture of the same body
        for f in own {
            let fx = f.borrow();
            let shape = fx.m_shape.as_ref().unwrap();
            let fi = fixtures.len() as i32;
            let shape_first = shapes.len() as i32;
            shape_recs(&**shape, &mut shapes);
            let proxy_first = if fx.m_proxy_count > 0 { proxies.len() as i32 } else { -1 };
            for p in fx.m_proxies.iter().take(fx.m_proxy_count as usize) {
                let pr = p.borrow();
                proxies.push(b2gpu_proxy_rec { fixture: fi, child_index: pr.child_index, proxy_id: pr.proxy_id, reserved: 0,
                                               aabb: [pr.aabb.lower_bound.x, pr.aabb.lower_bound.y, pr.aabb.upper_bound.x, pr.aabb.upper_bound.y] });
                proxy_ptrs.push(p.clone());
            }
            fixtures.push(b2gpu_fixture_rec {
                body: bi as i32, next: prev, shape_type: shape.get_type() as i32, shape_first,
                child_count: shape.get_child_count() as i32, proxy_first,
                density: fx.m_density, friction: fx.m_friction, restitution: fx.m_restitution,
                restitution_threshold: fx.m_restitution_threshold,
                category_bits: fx.m_filter.category_bits, mask_bits: fx.m_filter.mask_bits, group_index: fx.m_filter.group_index,
                is_sensor: fx.m_is_sensor as u16,
            });
            fixture_ptrs.push(f.clone());
            prev = fi;
        }
        fixture_head[bi] = prev; // newest fixture = list head
    }
    let fixture_index: HashMap<usize, i32> = fixture_ptrs.iter().enumerate().map(|(i, f)| (addr(f), i as i32)).collect();
    let proxy_index: HashMap<usize, i32> = proxy_ptrs.iter().enumerate().map(|(i, p)| (addr(p), i as i32)).collect();

    let bodies: Vec<b2gpu_body_rec> = body_ptrs.iter().enumerate().map(|(i, b)| {
        let b = b.borrow();
        b2gpu_body_rec {
            type_: b.m_type as i32, flags: b.m_flags.bits() as u32,
            xf_px: b.m_xf.p.x, xf_py: b.m_xf.p.y, xf_qs: b.m_xf.q.s, xf_qc: b.m_xf.q.c,
            lc_x: b.m_sweep.local_center.x, lc_y: b.m_sweep.local_center.y,
            c0_x: b.m_sweep.c0.x, c0_y: b.m_sweep.c0.y, c_x: b.m_sweep.c.x, c_y: b.m_sweep.c.y, a0: b.m_sweep.a0, a: b.m_sweep.a,
            vx: b.m_linear_velocity.x, vy: b.m_linear_velocity.y, w: b.m_angular_velocity,
            fx: b.m_force.x, fy: b.m_force.y, torque: b.m_torque,
            mass: b.m_mass, inv_mass: b.m_inv_mass, inertia: b.m_i, inv_inertia: b.m_inv_i,
            linear_damping: b.m_linear_damping, angular_damping: b.m_angular_damping, gravity_scale: b.m_gravity_scale,
            sleep_time: b.m_sleep_time, fixture_head: fixture_head[i], fixture_count: b.m_fixture_count, reserved: [0; 2],
        }
    }).collect();

    // ---- tree pool verbatim (B2dynamicTree::m_nodes, src/b2_dynamic_tree.rs:11-32), free nodes included
    let nodes: Vec<b2gpu_tree_node_rec> = tree.m_nodes.iter().take(tree.m_node_capacity as usize).map(|n| b2gpu_tree_node_rec {
        aabb: [n.aabb.lower_bound.x, n.aabb.lower_bound.y, n.aabb.upper_bound.x, n.aabb.upper_bound.y],
        parent: n.parent, child1: n.child1, child2: n.child2, height: n.height,
        proxy: if n.height == 0 { n.user_data.as_ref().map(|p| proxy_index[&addr(p)]).unwrap_or(-1) } else { -1 },
        moved: n.moved as i32,
    }).collect();

    // ---- contacts (world list reversed)
    let mut contact_ptrs: Vec<ContactPtr<D>> = cm.m_contact_list.iter().collect();
    contact_ptrs.reverse();
    let contacts: Vec<b2gpu_contact_rec> = contact_ptrs.iter().map(|c| {
        let c = c.borrow();
        let c = c.get_base();
        let m = &c.m_manifold;
        let pt = |i: usize| b2gpu_manifold_point {
            lp_x: m.points[i].local_point.x, lp_y: m.points[i].local_point.y,
            normal_impulse: m.points[i].normal_impulse, tangent_impulse: m.points[i].tangent_impulse, id: contact_id_key(&m.points[i].id),
        };
        b2gpu_contact_rec {
            fixture_a: fixture_index[&addr(&c.m_fixture_a)], fixture_b: fixture_index[&addr(&c.m_fixture_b)],
            index_a: c.m_index_a, index_b: c.m_index_b, flags: c.m_flags.bits() as u32,
            friction: c.m_friction, restitution: c.m_restitution, restitution_threshold: c.m_restitution_threshold,
            tangent_speed: c.m_tangent_speed, reserved: 0,
            manifold: b2gpu_manifold { points: [pt(0), pt(1)], ln_x: m.local_normal.x, ln_y: m.local_normal.y,
                                       lp_x: m.local_point.x, lp_y: m.local_point.y, type_: m.manifold_type as i32,
                                       point_count: m.point_count as i32 },
        }
    }).collect();

    // ---- joints (world list reversed); all ten joint types are inside the accelerated path
    let mut joint_ptrs: Vec<B2jointPtr<D>> = world.m_joint_list.iter().collect();
    joint_ptrs.reverse();
    let joints: Vec<b2gpu_joint_rec> = joint_ptrs.iter().map(|j| {
        let j = j.borrow();
        let base = j.get_base();
        let mut r = b2gpu_joint_rec::default();
        r.body_a = body_index[&addr(&base.m_body_a)];
        r.body_b = body_index[&addr(&base.m_body_b)];
        r.flags = if base.m_collide_connected { 1 } else { 0 };
        match j.as_derived() {
            JointAsDerived::ERevoluteJoint(v) => {
                r.type_ = 8;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_reference_angle; r.param[1] = v.m_lower_angle; r.param[2] = v.m_upper_angle;
                r.param[3] = v.m_max_motor_torque; r.param[4] = v.m_motor_speed;
                if v.m_enable_limit { r.flags |= 2; }
                if v.m_enable_motor { r.flags |= 4; }
                r.impulse[0] = v.m_impulse.x; r.impulse[1] = v.m_impulse.y; r.impulse[2] = v.m_motor_impulse;
                r.impulse[3] = v.m_lower_impulse; r.impulse[4] = v.m_upper_impulse;
            }
            JointAsDerived::EDistanceJoint(v) => {
                r.type_ = 1;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_length; r.param[1] = v.m_min_length; r.param[2] = v.m_max_length;
                r.param[3] = v.m_stiffness; r.param[4] = v.m_damping;
                r.impulse[0] = v.m_impulse; r.impulse[3] = v.m_lower_impulse; r.impulse[4] = v.m_upper_impulse;
            }
            JointAsDerived::EPrismaticJoint(v) => {
                r.type_ = 6;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_reference_angle; r.param[1] = v.m_lower_translation; r.param[2] = v.m_upper_translation;
                r.param[3] = v.m_max_motor_force; r.param[4] = v.m_motor_speed;
                r.param[5] = v.m_local_xaxis_a.x; r.param[6] = v.m_local_xaxis_a.y;
                if v.m_enable_limit { r.flags |= 2; }
                if v.m_enable_motor { r.flags |= 4; }
                r.impulse[0] = v.m_impulse.x; r.impulse[1] = v.m_impulse.y; r.impulse[2] = v.m_motor_impulse;
                r.impulse[3] = v.m_lower_impulse; r.impulse[4] = v.m_upper_impulse;
            }
            JointAsDerived::EFrictionJoint(v) => {
                r.type_ = 2;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_max_force; r.param[1] = v.m_max_torque;
                r.impulse[0] = v.m_linear_impulse.x; r.impulse[1] = v.m_linear_impulse.y; r.impulse[2] = v.m_angular_impulse;
            }
            JointAsDerived::EMotorJoint(v) => {
                r.type_ = 4;
                r.local_anchor_a = [v.m_linear_offset.x, v.m_linear_offset.y];
                r.param[0] = v.m_max_force; r.param[1] = v.m_max_torque;
                r.param[2] = v.m_angular_offset; r.param[3] = v.m_correction_factor;
                r.impulse[0] = v.m_linear_impulse.x; r.impulse[1] = v.m_linear_impulse.y; r.impulse[2] = v.m_angular_impulse;
            }
            JointAsDerived::EGearJoint(v) => {
                // four bodies: C / D and the static rest of the record overflow into impulse[1..6] (include/b2gpu.h)
                r.type_ = 3;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param = [v.m_local_anchor_c.x, v.m_local_anchor_c.y, v.m_local_anchor_d.x, v.m_local_anchor_d.y,
                           v.m_local_axis_c.x, v.m_local_axis_c.y, v.m_local_axis_d.x, v.m_local_axis_d.y];
                if v.m_type_a == B2jointType::EPrismaticJoint { r.flags |= 0x100; }
                if v.m_type_b == B2jointType::EPrismaticJoint { r.flags |= 0x200; }
                r.impulse[0] = v.m_impulse;
                r.impulse[1] = v.m_reference_angle_a; r.impulse[2] = v.m_reference_angle_b;
                r.impulse[3] = v.m_constant; r.impulse[4] = v.m_ratio;
                r.impulse[5] = f32::from_bits(body_index[&addr(&v.m_body_c)] as u32);
                r.impulse[6] = f32::from_bits(body_index[&addr(&v.m_body_d)] as u32);
            }
            JointAsDerived::EPulleyJoint(v) => {
                r.type_ = 7;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_ground_anchor_a.x; r.param[1] = v.m_ground_anchor_a.y;
                r.param[2] = v.m_ground_anchor_b.x; r.param[3] = v.m_ground_anchor_b.y;
                r.param[4] = v.m_length_a; r.param[5] = v.m_length_b; r.param[6] = v.m_ratio; r.param[7] = v.m_constant;
                r.impulse[0] = v.m_impulse;
            }
            JointAsDerived::EMouseJoint(v) => {
                r.type_ = 5;
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_max_force; r.param[1] = v.m_stiffness; r.param[2] = v.m_damping;
                r.param[3] = v.m_target_a.x; r.param[4] = v.m_target_a.y;
                r.impulse[0] = v.m_impulse.x; r.impulse[1] = v.m_impulse.y;
            }
            JointAsDerived::EWheelJoint(v) => {
                r.type_ = 10;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_stiffness; r.param[1] = v.m_lower_translation; r.param[2] = v.m_upper_translation;
                r.param[3] = v.m_max_motor_torque; r.param[4] = v.m_motor_speed;
                r.param[5] = v.m_local_xaxis_a.x; r.param[6] = v.m_local_xaxis_a.y; r.param[7] = v.m_damping;
                if v.m_enable_limit { r.flags |= 2; }
                if v.m_enable_motor { r.flags |= 4; }
                r.impulse[0] = v.m_impulse; r.impulse[1] = v.m_spring_impulse; r.impulse[2] = v.m_motor_impulse;
                r.impulse[3] = v.m_lower_impulse; r.impulse[4] = v.m_upper_impulse;
            }
            JointAsDerived::EWeldJoint(v) => {
                r.type_ = 9;
                r.local_anchor_a = [v.m_local_anchor_a.x, v.m_local_anchor_a.y];
                r.local_anchor_b = [v.m_local_anchor_b.x, v.m_local_anchor_b.y];
                r.param[0] = v.m_reference_angle; r.param[3] = v.m_stiffness; r.param[4] = v.m_damping;
                r.impulse[0] = v.m_impulse.x; r.impulse[1] = v.m_impulse.y; r.impulse[2] = v.m_impulse.z;
            }
            #[allow(unreachable_patterns)]
            _ => { r.type_ = 0; } // an unknown joint type: b2gpu_world_upload answers B2GPU_E_UNSUPPORTED
        }
        r
    }).collect();

    let mut flags: u32 = 0;
    if world.m_allow_sleep { flags |= 0x01; }
    if world.m_warm_starting { flags |= 0x02; }
    if world.m_new_contacts { flags |= 0x04; }
    if world.m_clear_forces { flags |= 0x08; }
    if crate::b2_contact::G_BLOCK_SOLVE.load(std::sync::atomic::Ordering::SeqCst) { flags |= 0x10; } // src/b2_contact.rs:25
    let wrec = b2gpu_world_rec {
        gravity_x: world.m_gravity.x, gravity_y: world.m_gravity.y, inv_dt0: world.m_inv_dt0, flags,
        tree_root: tree.m_root, tree_free_list: tree.m_free_list, tree_node_count: tree.m_node_count,
        tree_node_capacity: tree.m_node_capacity, tree_insertion_count: tree.m_insertion_count,
        proxy_count: bp.m_proxy_count, reserved: [0; 2],
    };
    let move_buffer: Vec<i32> = bp.m_move_buffer.iter().take(bp.m_move_count as usize).cloned().collect();
    Snapshot { world: wrec, bodies, fixtures, shapes, proxies, nodes, contacts, move_buffer, joints,
               body_ptrs, fixture_ptrs, proxy_ptrs, joint_ptrs }
}

impl<D: UserDataType> Snapshot<D> {
    pub fn sizes(&self) -> b2gpu_snapshot_sizes {
        b2gpu_snapshot_sizes {
            body_count: self.bodies.len() as i32, fixture_count: self.fixtures.len() as i32, shape_count: self.shapes.len() as i32,
            proxy_count: self.proxies.len() as i32, node_count: self.nodes.len() as i32, contact_count: self.contacts.len() as i32,
            move_count: self.move_buffer.len() as i32, joint_count: self.joints.len() as i32,
        }
    }
    /// The C view of the tables (valid while `self` is alive and unmodified).
    pub fn as_raw(&mut self) -> b2gpu_snapshot {
        b2gpu_snapshot {
            world: self.world, n: self.sizes(),
            bodies: self.bodies.as_mut_ptr(), fixtures: self.fixtures.as_mut_ptr(), shapes: self.shapes.as_mut_ptr(),
            proxies: self.proxies.as_mut_ptr(), nodes: self.nodes.as_mut_ptr(), contacts: self.contacts.as_mut_ptr(),
            move_buffer: self.move_buffer.as_mut_ptr(), joints: self.joints.as_mut_ptr(),
        }
    }
    /// Room for a download whose contact / move / node tables may have grown (sizes from b2gpu_world_snapshot_sizes).
    pub fn reserve_for(&mut self, n: &b2gpu_snapshot_sizes) {
        self.nodes.resize(n.node_count as usize, b2gpu_tree_node_rec::default());
        self.contacts.resize(n.contact_count as usize, b2gpu_contact_rec::default());
        self.move_buffer.resize(n.move_count as usize, -1);
    }

    /// The checkpoint file of box2d_rs_b200/csrc/b2g_checkpoint.cu (file version 2), written without the library so that
    /// `examples/dump_state.rs` needs neither nvcc nor a GPU.  Little-endian hosts only (the file carries a byte-order tag).
    pub fn save(&self, path: &std::path::Path) -> std::io::Result<()> {
        fn bytes_of<T: Copy>(v: &[T]) -> &[u8] {
            // plain-old-data records (#[repr(C)], no padding beyond what the header documents)
            unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u8, std::mem::size_of_val(v)) }
        }
        fn fnv1a(data: &[u8], mut h: u64) -> u64 {
            for b in data { h ^= *b as u64; h = h.wrapping_mul(1099511628211); }
            h
        }
        let tables: [&[u8]; 8] = [bytes_of(&self.bodies), bytes_of(&self.fixtures), bytes_of(&self.shapes), bytes_of(&self.proxies),
                                  bytes_of(&self.nodes), bytes_of(&self.contacts), bytes_of(&self.move_buffer), bytes_of(&self.joints)];
        let record_bytes: [u32; 8] = [128, 48, 160, 32, 40, 104, 4, 96];
        let mut payload_hash: u64 = 1469598103934665603;
        let mut payload_bytes: u64 = 0;
        for t in tables.iter() { payload_hash = fnv1a(t, payload_hash); payload_bytes += t.len() as u64; }
        let mut h: Vec<u8> = Vec::new();
        h.extend_from_slice(b"B2GPUSNP");
        for v in [2u32 /* file version */, B2GPU_ABI_VERSION as u32, 0x01020304u32, 0u32 /* header bytes, patched below */] {
            h.extend_from_slice(&v.to_le_bytes());
        }
        for v in record_bytes { h.extend_from_slice(&v.to_le_bytes()); }
        h.extend_from_slice(bytes_of(std::slice::from_ref(&self.world)));
        h.extend_from_slice(bytes_of(std::slice::from_ref(&self.sizes())));
        h.extend_from_slice(&payload_bytes.to_le_bytes());
        h.extend_from_slice(&payload_hash.to_le_bytes());
        let header_bytes = (h.len() + 8) as u32; // + header_hash; FileHeader has no padding: 8 + 16 + 32 + 48 + 32 + 24 = 160
        h[20..24].copy_from_slice(&header_bytes.to_le_bytes());
        let header_hash = fnv1a(&h, 1469598103934665603);
        h.extend_from_slice(&header_hash.to_le_bytes());
        let mut out = h;
        for t in tables.iter() { out.extend_from_slice(t); }
        let tmp = path.with_extension("tmp");
        std::fs::write(&tmp, &out)?;
        std::fs::rename(&tmp, path)
    }
}

/// Device results back into the world graph after `b2gpu_world_download(&mut snap.as_raw())`.  `snap` must be the
/// snapshot `flatten` produced for this world (it carries the Rc handles), with tables refreshed by the download.
pub fn write_back<D: UserDataType>(world: &mut B2world<D>, snap: &Snapshot<D>) {
    // ---- bodies: m_xf, m_sweep, velocities, forces, flags, sleep time (b2_island_private.rs:277-327)
    for (b, r) in snap.body_ptrs.iter().zip(snap.bodies.iter()) {
        let mut b = b.borrow_mut();
        b.m_flags = BodyFlags::from_bits_truncate(r.flags as u16);
        b.m_xf.p.set(r.xf_px, r.xf_py);
        b.m_xf.q.s = r.xf_qs;
        b.m_xf.q.c = r.xf_qc;
        b.m_sweep.c0.set(r.c0_x, r.c0_y);
        b.m_sweep.c.set(r.c_x, r.c_y);
        b.m_sweep.a0 = r.a0;
        b.m_sweep.a = r.a;
        b.m_linear_velocity.set(r.vx, r.vy);
        b.m_angular_velocity = r.w;
        b.m_force.set(r.fx, r.fy);
        b.m_torque = r.torque;
        b.m_sleep_time = r.sleep_time;
    }
    let cm_ptr = world.m_contact_manager.clone();
    {
        // ---- broadphase: proxies' tight boxes, the tree pool, the move buffer
        let cm = cm_ptr.borrow();
        let mut bp = cm.m_broad_phase.borrow_mut();
        for (p, r) in snap.proxy_ptrs.iter().zip(snap.proxies.iter()) {
            let mut p = p.borrow_mut();
            p.aabb.lower_bound.set(r.aabb[0], r.aabb[1]);
            p.aabb.upper_bound.set(r.aabb[2], r.aabb[3]);
        }
        let tree = &mut bp.m_tree;
        tree.m_nodes.resize(snap.nodes.len(), Default::default());
        for (n, r) in tree.m_nodes.iter_mut().zip(snap.nodes.iter()) {
            n.aabb.lower_bound.set(r.aabb[0], r.aabb[1]);
            n.aabb.upper_bound.set(r.aabb[2], r.aabb[3]);
            n.parent = r.parent; n.child1 = r.child1; n.child2 = r.child2; n.height = r.height;
            n.moved = r.moved != 0;
            n.user_data = if r.proxy >= 0 { Some(snap.proxy_ptrs[r.proxy as usize].clone()) } else { None };
        }
        tree.m_root = snap.world.tree_root;
        tree.m_free_list = snap.world.tree_free_list;
        tree.m_node_count = snap.world.tree_node_count;
        tree.m_node_capacity = snap.world.tree_node_capacity;
        tree.m_insertion_count = snap.world.tree_insertion_count;
        bp.m_move_buffer.clear();
        bp.m_move_buffer.extend_from_slice(&snap.move_buffer);
        bp.m_move_count = snap.move_buffer.len() as i32;
        bp.m_move_capacity = bp.m_move_capacity.max(bp.m_move_count);
    }
    // ---- contacts: the device keeps creation order with stable compaction, so the new list is (survivors of the old
    //      list, in order) followed by the contacts created in this step.  Destroy what vanished (the manager's own
    //      destroy: unlinks the edges, end_contact to the listener), update survivors, create the rest as add_pair's
    //      tail does (b2_contact_manager.rs(private):236-300).
    type Key = (i32, i32, i32, i32);
    let key_of = |c: &b2gpu_contact_rec| -> Key { (c.fixture_a, c.index_a, c.fixture_b, c.index_b) };
    let fixture_index: HashMap<usize, i32> = snap.fixture_ptrs.iter().enumerate().map(|(i, f)| (addr(f), i as i32)).collect();
    let wanted: HashMap<Key, usize> = snap.contacts.iter().enumerate().map(|(i, c)| (key_of(c), i)).collect();
    let mut existing: HashMap<Key, ContactPtr<D>> = HashMap::new();
    let old: Vec<ContactPtr<D>> = cm_ptr.borrow().m_contact_list.iter().collect();
    for c in old {
        let k = {
            let cb = c.borrow();
            let cb = cb.get_base();
            (fixture_index[&addr(&cb.m_fixture_a)], cb.m_index_a, fixture_index[&addr(&cb.m_fixture_b)], cb.m_index_b)
        };
        if wanted.contains_key(&k) { existing.insert(k, c); } else { cm_ptr.borrow_mut().destroy(c); }
    }
    for r in snap.contacts.iter() { // ascending = creation order: push_front leaves the newest at the head
        let k = key_of(r);
        let c = match existing.get(&k) {
            Some(c) => c.clone(),
            None => {
                let fa = snap.fixture_ptrs[r.fixture_a as usize].clone();
                let fb = snap.fixture_ptrs[r.fixture_b as usize].clone();
                // the record already carries the register-order swap, so create_fcn is called in primary order
                let c = B2contact::create(&*cm_ptr.borrow(), fa.clone(), r.index_a, fb.clone(), r.index_b);
                let (body_a, body_b) = (fa.borrow().get_body(), fb.borrow().get_body());
                let mut cm = cm_ptr.borrow_mut();
                cm.m_contact_list.push_front(c.clone());
                let node_a = Rc::new(RefCell::new(B2contactEdge { contact: Rc::downgrade(&c), other: Rc::downgrade(&body_b), prev: None, next: None }));
                c.borrow_mut().get_base_mut().m_node_a = Some(node_a.clone());
                body_a.borrow_mut().m_contact_list.push_front(node_a);
                let node_b = Rc::new(RefCell::new(B2contactEdge { contact: Rc::downgrade(&c), other: Rc::downgrade(&body_a), prev: None, next: None }));
                c.borrow_mut().get_base_mut().m_node_b = Some(node_b.clone());
                body_b.borrow_mut().m_contact_list.push_front(node_b);
                cm.m_contact_count += 1;
                c
            }
        };
        let mut cb = c.borrow_mut();
        let cb = cb.get_base_mut();
        cb.m_flags = ContactFlags::from_bits_truncate(r.flags);
        cb.m_friction = r.friction;
        cb.m_restitution = r.restitution;
        cb.m_restitution_threshold = r.restitution_threshold;
        cb.m_tangent_speed = r.tangent_speed;
        let m = &mut cb.m_manifold;
        for i in 0..2 {
            let p = &r.manifold.points[i];
            m.points[i].local_point.set(p.lp_x, p.lp_y);
            m.points[i].normal_impulse = p.normal_impulse;
            m.points[i].tangent_impulse = p.tangent_impulse;
            m.points[i].id = contact_id_from_key(p.id);
        }
        m.local_normal.set(r.manifold.ln_x, r.manifold.ln_y);
        m.local_point.set(r.manifold.lp_x, r.manifold.lp_y);
        m.manifold_type = match r.manifold.type_ { 0 => B2manifoldType::ECircles, 1 => B2manifoldType::EFaceA, _ => B2manifoldType::EFaceB };
        m.point_count = r.manifold.point_count as usize;
    }
    // NOTE: a world stepped on the device in exact mode (or large-world mode 2) keeps the survivors in their old
    // relative order, which is the order they already have in m_contact_list, so no relinking is needed for them.

    // ---- joints: accumulated impulses (warm start) and the per-world motor / limit switches
    for (j, r) in snap.joint_ptrs.iter().zip(snap.joints.iter()) {
        match j.borrow_mut().as_derived_mut() {
            JointAsDerivedMut::ERevoluteJoint(v) => {
                v.m_impulse.set(r.impulse[0], r.impulse[1]);
                v.m_motor_impulse = r.impulse[2];
                v.m_lower_impulse = r.impulse[3];
                v.m_upper_impulse = r.impulse[4];
            }
            JointAsDerivedMut::EDistanceJoint(v) => {
                v.m_impulse = r.impulse[0];
                v.m_lower_impulse = r.impulse[3];
                v.m_upper_impulse = r.impulse[4];
            }
            JointAsDerivedMut::EPrismaticJoint(v) => {
                v.m_impulse.set(r.impulse[0], r.impulse[1]);
                v.m_motor_impulse = r.impulse[2];
                v.m_lower_impulse = r.impulse[3];
                v.m_upper_impulse = r.impulse[4];
            }
            JointAsDerivedMut::EFrictionJoint(v) => {
                v.m_linear_impulse.set(r.impulse[0], r.impulse[1]);
                v.m_angular_impulse = r.impulse[2];
            }
            JointAsDerivedMut::EMotorJoint(v) => {
                v.m_linear_impulse.set(r.impulse[0], r.impulse[1]);
                v.m_angular_impulse = r.impulse[2];
            }
            JointAsDerivedMut::EPulleyJoint(v) => {
                v.m_impulse = r.impulse[0];
            }
            JointAsDerivedMut::EGearJoint(v) => {
                v.m_impulse = r.impulse[0];
            }
            JointAsDerivedMut::EMouseJoint(v) => {
                v.m_impulse.set(r.impulse[0], r.impulse[1]);
            }
            JointAsDerivedMut::EWheelJoint(v) => {
                v.m_impulse = r.impulse[0];
                v.m_spring_impulse = r.impulse[1];
                v.m_motor_impulse = r.impulse[2];
                v.m_lower_impulse = r.impulse[3];
                v.m_upper_impulse = r.impulse[4];
            }
            JointAsDerivedMut::EWeldJoint(v) => {
                v.m_impulse.set(r.impulse[0], r.impulse[1], r.impulse[2]);
            }
            _ => {}
        }
    }
    world.m_inv_dt0 = snap.world.inv_dt0;
    world.m_new_contacts = snap.world.flags & 0x04 != 0;
}
