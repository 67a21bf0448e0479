// TEST INFRASTRUCTURE — C API of the CPU oracle for ctypes (tests/, bench.py cpu_baseline /
// --impl reference, __graft_entry__.smoke only).  Never linked into the product.
// PARITY UNPINNED beyond the reference's own tests (see b2o_math.hpp).
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/b2gpu.h"
#include "b2o_world.hpp"
#include "b2o_query.hpp"

using namespace b2o;

static Shape shape_from_def(const b2gpu_shape_def* d) {
  Shape s;
  s.type = d->type;
  s.radius = d->radius;
  s.p = Vec2(d->p_x, d->p_y);
  s.v0 = Vec2(d->v0[0], d->v0[1]);
  s.v1 = Vec2(d->v1[0], d->v1[1]);
  s.v2 = Vec2(d->v2[0], d->v2[1]);
  s.v3 = Vec2(d->v3[0], d->v3[1]);
  s.one_sided = d->one_sided != 0;
  s.count = d->count;
  s.centroid = Vec2(d->centroid[0], d->centroid[1]);
  for (int i = 0; i < MAX_POLYGON_VERTICES; ++i) {
    s.vertices[i] = Vec2(d->vertices[2 * i], d->vertices[2 * i + 1]);
    s.normals[i] = Vec2(d->normals[2 * i], d->normals[2 * i + 1]);
  }
  if (d->type == E_CHAIN) {
    for (int i = 0; i < d->chain_count; ++i) s.chain.push_back(Vec2(d->chain_vertices[2 * i], d->chain_vertices[2 * i + 1]));
    s.chain_prev = Vec2(d->chain_prev[0], d->chain_prev[1]);
    s.chain_next = Vec2(d->chain_next[0], d->chain_next[1]);
  }
  return s;
}
static void shape_to_def(const Shape& s, b2gpu_shape_def* d) {
  d->type = s.type;
  d->radius = s.radius;
  d->p_x = s.p.x; d->p_y = s.p.y;
  d->count = s.count;
  d->centroid[0] = s.centroid.x; d->centroid[1] = s.centroid.y;
  for (int i = 0; i < MAX_POLYGON_VERTICES; ++i) {
    d->vertices[2 * i] = s.vertices[i].x; d->vertices[2 * i + 1] = s.vertices[i].y;
    d->normals[2 * i] = s.normals[i].x; d->normals[2 * i + 1] = s.normals[i].y;
  }
}

extern "C" {

// ---- shapes (restated setup geometry, used to pin collision_test.rs)
int b2o_polygon_set_as_box(b2gpu_shape_def* d, float hx, float hy) {
  Shape s; polygon_set_as_box(s, hx, hy); shape_to_def(s, d); return 0;
}
int b2o_polygon_set_as_box_angle(b2gpu_shape_def* d, float hx, float hy, float cx, float cy, float angle) {
  Shape s; polygon_set_as_box_angle(s, hx, hy, Vec2(cx, cy), angle); shape_to_def(s, d); return 0;
}
int b2o_polygon_set(b2gpu_shape_def* d, const float* xy, int count) {
  std::vector<Vec2> v;
  for (int i = 0; i < count; ++i) v.push_back(Vec2(xy[2 * i], xy[2 * i + 1]));
  Shape s; bool ok = polygon_set(s, v.data(), count); shape_to_def(s, d); return ok ? 0 : -1;
}
int b2o_shape_compute_mass(const b2gpu_shape_def* d, float density, b2gpu_mass_data* out) {
  Shape s = shape_from_def(d);
  MassData md; shape_compute_mass(s, md, density);
  out->mass = md.mass; out->center_x = md.center.x; out->center_y = md.center.y; out->inertia = md.i;
  return 0;
}
// b2_distance_fn between two shapes' children under (p.x, p.y, angle) transforms: out5 = point_a, point_b, distance
int b2o_shape_distance(const b2gpu_shape_def* da, int index_a, const float* xa3, const b2gpu_shape_def* db, int index_b,
                       const float* xb3, int use_radii, float* out5) {
  Shape a = shape_from_def(da), b = shape_from_def(db);
  DistanceProxy pa, pb;
  proxy_set_shape(pa, a, index_a);
  proxy_set_shape(pb, b, index_b);
  Transform ta, tb;
  ta.p = Vec2(xa3[0], xa3[1]); ta.q.set(xa3[2]);
  tb.p = Vec2(xb3[0], xb3[1]); tb.q.set(xb3[2]);
  SimplexCache cache;
  DistanceOutput o;
  b2_distance(o, cache, pa, ta, pb, tb, use_radii != 0);
  out5[0] = o.point_a.x; out5[1] = o.point_a.y; out5[2] = o.point_b.x; out5[3] = o.point_b.y; out5[4] = o.distance;
  return o.iterations;
}
int b2o_test_overlap_shapes(const b2gpu_shape_def* da, int index_a, const float* xa3, const b2gpu_shape_def* db, int index_b,
                            const float* xb3) {
  Shape a = shape_from_def(da), b = shape_from_def(db);
  Transform ta, tb;
  ta.p = Vec2(xa3[0], xa3[1]); ta.q.set(xa3[2]);
  tb.p = Vec2(xb3[0], xb3[1]); tb.q.set(xb3[2]);
  return b2_test_overlap_shapes(a, index_a, b, index_b, ta, tb) ? 1 : 0;
}
// tests/math_test.rs:25-49 — sweep endpoints
void b2o_sweep_get_transform(const float* sweep8 /*lc c0 c a0 a*/, float beta, float* xf4) {
  Vec2 lc(sweep8[0], sweep8[1]), c0(sweep8[2], sweep8[3]), c(sweep8[4], sweep8[5]);
  float a0 = sweep8[6], a = sweep8[7];
  Vec2 p = (1.0f - beta) * c0 + beta * c;  // src/b2_math.rs:762-769
  float angle = (1.0f - beta) * a0 + beta * a;
  Rot q(angle);
  p -= b2_mul_rot(q, lc);
  xf4[0] = p.x; xf4[1] = p.y; xf4[2] = q.s; xf4[3] = q.c;
}

// ---- world
void* b2o_world_create(float gx, float gy) { return new World(Vec2(gx, gy)); }
void b2o_world_destroy(void* w) { delete (World*)w; }
void* b2o_world_clone(void* w) { return new World(*(World*)w); }
int b2o_create_body(void* w, const b2gpu_body_def* d) {
  BodyDef bd;
  bd.type = d->type;
  bd.position = Vec2(d->position_x, d->position_y);
  bd.angle = d->angle;
  bd.linear_velocity = Vec2(d->linear_velocity_x, d->linear_velocity_y);
  bd.angular_velocity = d->angular_velocity;
  bd.linear_damping = d->linear_damping;
  bd.angular_damping = d->angular_damping;
  bd.allow_sleep = d->allow_sleep; bd.awake = d->awake; bd.fixed_rotation = d->fixed_rotation;
  bd.bullet = d->bullet; bd.enabled = d->enabled;
  bd.gravity_scale = d->gravity_scale;
  return ((World*)w)->create_body(bd);
}
int b2o_create_fixture(void* w, int body, const b2gpu_fixture_def* d, const b2gpu_shape_def* sd) {
  FixtureDef fd;
  fd.friction = d->friction; fd.restitution = d->restitution; fd.restitution_threshold = d->restitution_threshold;
  fd.density = d->density; fd.is_sensor = d->is_sensor != 0;
  fd.filter.category_bits = d->category_bits; fd.filter.mask_bits = d->mask_bits; fd.filter.group_index = d->group_index;
  return ((World*)w)->create_fixture(body, fd, shape_from_def(sd));
}
void b2o_set_transform(void* w, int body, float px, float py, float angle) { ((World*)w)->set_transform(body, Vec2(px, py), angle); }
void b2o_set_linear_velocity(void* w, int body, float vx, float vy) { ((World*)w)->set_linear_velocity(body, Vec2(vx, vy)); }
void b2o_set_angular_velocity(void* w, int body, float av) { ((World*)w)->set_angular_velocity(body, av); }
void b2o_apply_force_to_center(void* w, int body, float fx, float fy, int wake) { ((World*)w)->apply_force_to_center(body, Vec2(fx, fy), wake != 0); }
void b2o_apply_force(void* w, int body, float fx, float fy, float px, float py, int wake) { ((World*)w)->apply_force(body, Vec2(fx, fy), Vec2(px, py), wake != 0); }
void b2o_apply_torque(void* w, int body, float t, int wake) { ((World*)w)->apply_torque(body, t, wake != 0); }
void b2o_apply_linear_impulse(void* w, int body, float ix, float iy, float px, float py, int wake) { ((World*)w)->apply_linear_impulse(body, Vec2(ix, iy), Vec2(px, py), wake != 0); }
void b2o_apply_linear_impulse_to_center(void* w, int body, float ix, float iy, int wake) { ((World*)w)->apply_linear_impulse_to_center(body, Vec2(ix, iy), wake != 0); }
void b2o_apply_angular_impulse(void* w, int body, float i, int wake) { ((World*)w)->apply_angular_impulse(body, i, wake != 0); }
void b2o_body_set_awake(void* w, int body, int flag) { ((World*)w)->set_awake(body, flag != 0); }
void b2o_body_set_damping(void* w, int body, float l, float a) { Body& b = ((World*)w)->bodies[body]; b.linear_damping = l; b.angular_damping = a; }  // src/b2_body.rs:755-765
void b2o_body_set_gravity_scale(void* w, int body, float s) { ((World*)w)->bodies[body].gravity_scale = s; }  // :771-773
void b2o_body_set_sleeping_allowed(void* w, int body, int flag) {  // :815-821
  World* W = (World*)w;
  if (flag) W->bodies[body].flags |= BF_AUTO_SLEEP;
  else { W->bodies[body].flags &= ~BF_AUTO_SLEEP; W->set_awake(body, true); }
}
// ---- joints (b2gpu_joint_def in and out, so the scene recipes drive both engines)
static void joint_def_out(const JointDef& d, b2gpu_joint_def* o) {
  std::memset(o, 0, sizeof(*o));
  o->type = d.type; o->body_a = d.body_a; o->body_b = d.body_b; o->collide_connected = d.collide_connected ? 1 : 0;
  o->local_anchor_a[0] = d.local_anchor_a.x; o->local_anchor_a[1] = d.local_anchor_a.y;
  o->local_anchor_b[0] = d.local_anchor_b.x; o->local_anchor_b[1] = d.local_anchor_b.y;
  o->reference_angle = d.reference_angle; o->lower_angle = d.lower_angle; o->upper_angle = d.upper_angle;
  o->max_motor_torque = d.max_motor_torque; o->motor_speed = d.motor_speed;
  o->enable_limit = d.enable_limit ? 1 : 0; o->enable_motor = d.enable_motor ? 1 : 0;
  o->length = d.length; o->min_length = d.min_length; o->max_length = d.max_length; o->stiffness = d.stiffness; o->damping = d.damping;
  if (d.type == J_FRICTION || d.type == J_MOTOR) {  // b2gpu.h: length = max_force, stiffness = correction_factor (motor)
    o->length = d.max_force; o->min_length = 0.0f; o->max_length = 0.0f;
    o->stiffness = d.type == J_MOTOR ? d.correction_factor : 0.0f;
  }
  if (d.type == J_PULLEY) {  // b2gpu.h: ground anchors over (lower_angle, upper_angle) / (max_motor_torque, motor_speed)
    o->lower_angle = d.ground_anchor_a.x; o->upper_angle = d.ground_anchor_a.y;
    o->max_motor_torque = d.ground_anchor_b.x; o->motor_speed = d.ground_anchor_b.y;
    o->length = d.length; o->min_length = d.length_b; o->max_length = d.ratio;
  }
  if (d.type == J_GEAR) {  // b2gpu.h: enable_limit / enable_motor = joint1 / joint2 (indices), length = ratio
    o->enable_limit = d.joint1; o->enable_motor = d.joint2;
    o->length = d.ratio; o->min_length = 0.0f; o->max_length = 0.0f;
  }
  if (d.type == J_MOUSE) {  // b2gpu.h: local_anchor_a = target (world), length = max_force
    o->local_anchor_a[0] = d.target.x; o->local_anchor_a[1] = d.target.y;
    o->length = d.max_force; o->min_length = 0.0f; o->max_length = 0.0f;
  }
  if (d.type == J_PRISMATIC || d.type == J_WHEEL) { o->length = d.local_axis_a.x; o->min_length = d.local_axis_a.y; o->max_length = 0.0f; }  // b2gpu.h: the def's overlay
}
int b2o_prismatic_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay, float dx, float dy) {
  joint_def_out(((World*)w)->prismatic_joint_def(body_a, body_b, Vec2(ax, ay), Vec2(dx, dy)), def);
  return 0;
}
int b2o_revolute_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay) {
  joint_def_out(((World*)w)->revolute_joint_def(body_a, body_b, Vec2(ax, ay)), def);
  return 0;
}
int b2o_distance_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float a1x, float a1y, float a2x, float a2y) {
  joint_def_out(((World*)w)->distance_joint_def(body_a, body_b, Vec2(a1x, a1y), Vec2(a2x, a2y)), def);
  return 0;
}
int b2o_friction_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay) {
  joint_def_out(((World*)w)->friction_joint_def(body_a, body_b, Vec2(ax, ay)), def);
  return 0;
}
int b2o_pulley_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float gax, float gay, float gbx, float gby,
                         float ax, float ay, float bx, float by, float ratio) {
  joint_def_out(((World*)w)->pulley_joint_def(body_a, body_b, Vec2(gax, gay), Vec2(gbx, gby), Vec2(ax, ay), Vec2(bx, by), ratio), def);
  return 0;
}
int b2o_gear_joint_def(void* w, b2gpu_joint_def* def, int joint1, int joint2, float ratio) {
  joint_def_out(((World*)w)->gear_joint_def(joint1, joint2, ratio), def);
  return 0;
}
int b2o_mouse_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float tx, float ty) {
  joint_def_out(((World*)w)->mouse_joint_def(body_a, body_b, Vec2(tx, ty)), def);
  return 0;
}
int b2o_motor_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b) {
  joint_def_out(((World*)w)->motor_joint_def(body_a, body_b), def);
  return 0;
}
int b2o_wheel_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay, float dx, float dy) {
  joint_def_out(((World*)w)->wheel_joint_def(body_a, body_b, Vec2(ax, ay), Vec2(dx, dy)), def);
  return 0;
}
int b2o_weld_joint_def(void* w, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay) {
  joint_def_out(((World*)w)->weld_joint_def(body_a, body_b, Vec2(ax, ay)), def);
  return 0;
}
int b2o_angular_stiffness(void* w, float hz, float ratio, int body_a, int body_b, float* stiffness, float* damping) {
  ((World*)w)->angular_stiffness(*stiffness, *damping, hz, ratio, body_a, body_b);
  return 0;
}
int b2o_linear_stiffness(void* w, float hz, float ratio, int body_a, int body_b, float* stiffness, float* damping) {
  ((World*)w)->linear_stiffness(*stiffness, *damping, hz, ratio, body_a, body_b);
  return 0;
}
int b2o_create_joint(void* w, const b2gpu_joint_def* d) {
  JointDef jd;
  jd.type = d->type; jd.body_a = d->body_a; jd.body_b = d->body_b; jd.collide_connected = d->collide_connected != 0;
  jd.local_anchor_a = Vec2(d->local_anchor_a[0], d->local_anchor_a[1]);
  jd.local_anchor_b = Vec2(d->local_anchor_b[0], d->local_anchor_b[1]);
  jd.reference_angle = d->reference_angle; jd.lower_angle = d->lower_angle; jd.upper_angle = d->upper_angle;
  jd.max_motor_torque = d->max_motor_torque; jd.motor_speed = d->motor_speed;
  jd.enable_limit = d->enable_limit != 0; jd.enable_motor = d->enable_motor != 0;
  jd.length = d->length; jd.min_length = d->min_length; jd.max_length = d->max_length; jd.stiffness = d->stiffness; jd.damping = d->damping;
  if (d->type == J_PRISMATIC || d->type == J_WHEEL) jd.local_axis_a = Vec2(d->length, d->min_length);
  if (d->type == J_FRICTION || d->type == J_MOTOR) { jd.max_force = d->length; jd.correction_factor = d->stiffness; }
  if (d->type == J_PULLEY) {
    jd.ground_anchor_a = Vec2(d->lower_angle, d->upper_angle); jd.ground_anchor_b = Vec2(d->max_motor_torque, d->motor_speed);
    jd.length_b = d->min_length; jd.ratio = d->max_length;
  }
  if (d->type == J_MOUSE) { jd.target = jd.local_anchor_a; jd.max_force = d->length; }
  if (d->type == J_GEAR) { jd.joint1 = d->enable_limit; jd.joint2 = d->enable_motor; jd.ratio = d->length; jd.enable_limit = jd.enable_motor = false; }
  return ((World*)w)->create_joint(jd);
}
void b2o_destroy_joint(void* w, int j) { ((World*)w)->destroy_joint(j); }
int b2o_joint_count(void* w) { return (int)((World*)w)->joints.size(); }
void b2o_joint_set_target(void* w, int j, float x, float y) { ((World*)w)->joint_set_target(j, Vec2(x, y)); }
void b2o_joint_set_motor_speed(void* w, int j, float v) { ((World*)w)->joint_set_motor_speed(j, v); }
void b2o_joint_set_max_motor_torque(void* w, int j, float v) { ((World*)w)->joint_set_max_motor_torque(j, v); }
void b2o_joint_enable_motor(void* w, int j, int f) { ((World*)w)->joint_enable_motor(j, f != 0); }
void b2o_joint_enable_limit(void* w, int j, int f) { ((World*)w)->joint_enable_limit(j, f != 0); }
void b2o_joint_set_limits(void* w, int j, float lo, float hi) { ((World*)w)->joint_set_limits(j, lo, hi); }
void b2o_set_allow_sleeping(void* w, int f) {  // b2_world.rs(private):340-353
  World* W = (World*)w;
  if ((f != 0) == W->allow_sleep) return;
  W->allow_sleep = f != 0;
  if (!W->allow_sleep)
    for (int b = W->body_list; b != -1; b = W->bodies[b].next) W->set_awake(b, true);
}
void b2o_set_gravity(void* w, float gx, float gy) { ((World*)w)->gravity = Vec2(gx, gy); }  // src/b2_world.rs:358-360
void b2o_set_warm_starting(void* w, int f) { ((World*)w)->warm_starting = f != 0; }
void b2o_set_block_solve(void* w, int f) { ((World*)w)->block_solve = f != 0; }
void b2o_set_collect_levels(void* w, int f) { ((World*)w)->collect_levels = f != 0; }
// largest island of the last step: out10 = contacts, bodies, sweeps, depth, depth of one sweep, handover, makespan x4
void b2o_set_collect_dag(void* w, int f, double handover) { World* W = (World*)w; W->collect_dag = f != 0; W->collect_levels = W->collect_levels || f != 0; W->dag_handover = handover; }
void b2o_get_dag_stats(void* w, double* out10) {
  const World::DagStats& d = ((World*)w)->dag;
  out10[0] = d.contacts; out10[1] = d.bodies; out10[2] = d.sweeps; out10[3] = d.depth; out10[4] = d.depth_one_sweep;
  out10[5] = ((World*)w)->dag_handover;
  for (int i = 0; i < 4; ++i) out10[6 + i] = d.makespan[i];
}
void b2o_get_dag_cyclic(void* w, double* out12) {
  const World::DagStats& d = ((World*)w)->dag;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) out12[3 * i + j] = d.makespan_cyclic[i][j];
}
void b2o_step(void* w, float dt, int vi, int pi) { ((World*)w)->step(dt, vi, pi); }
int b2o_body_count(void* w) { return (int)((World*)w)->bodies.size(); }
int b2o_contact_count(void* w) { return ((World*)w)->contact_count; }
void b2o_get_profile(void* w, double* out7) {
  const Profile& p = ((World*)w)->profile;
  out7[0] = p.step; out7[1] = p.collide; out7[2] = p.solve; out7[3] = p.solve_init; out7[4] = p.solve_velocity;
  out7[5] = p.solve_position; out7[6] = p.broadphase;
}
// begin/end contact events of the last step in firing order: out = [cap][5] (type, fixture_a, index_a, fixture_b,
// index_b); returns the number of events (may exceed cap).
int b2o_get_events(void* w, int32_t* out, int cap) {
  const std::vector<ContactEvent>& ev = ((World*)w)->events;
  for (size_t i = 0; i < ev.size() && (int)i < cap; ++i) {
    out[5 * i] = ev[i].type; out[5 * i + 1] = ev[i].fixture_a; out[5 * i + 2] = ev[i].index_a;
    out[5 * i + 3] = ev[i].fixture_b; out[5 * i + 4] = ev[i].index_b;
  }
  return (int)ev.size();
}
// post_solve reports of the last step in call order: out = b2gpu_post_solve_event[cap]; returns the number (may exceed cap)
int b2o_get_post_solve(void* w, b2gpu_post_solve_event* out, int cap) {
  const std::vector<PostSolveEvent>& ev = ((World*)w)->post_solve_events;
  for (size_t i = 0; i < ev.size() && (int)i < cap; ++i) {
    std::memset(&out[i], 0, sizeof(out[i]));
    out[i].fixture_a = ev[i].fixture_a; out[i].index_a = ev[i].index_a; out[i].fixture_b = ev[i].fixture_b; out[i].index_b = ev[i].index_b;
    out[i].count = ev[i].count;
    for (int j = 0; j < 2; ++j) { out[i].normal_impulses[j] = ev[i].normal_impulses[j]; out[i].tangent_impulses[j] = ev[i].tangent_impulses[j]; }
  }
  return (int)ev.size();
}
void b2o_get_stats(void* w, b2gpu_step_stats* out) {
  const StepStats& s = ((World*)w)->stats;
  std::memset(out, 0, sizeof(*out));
  out->contacts = s.contacts; out->touching = s.touching; out->destroyed = s.destroyed; out->islands = s.islands;
  out->island_bodies = s.island_bodies; out->island_contacts = s.island_contacts; out->moved = s.moved; out->pairs = s.pairs;
  out->created = s.created; out->awake_bodies = s.awake_bodies; out->solver_levels = s.solver_levels;
}

// ---- world queries (b2o_query.hpp): same record layout as b2gpu_ray_hit / b2gpu_world_query_aabb
void b2o_ray_cast_closest(void* w, const float* p1p2, int n, b2gpu_ray_hit* out) {
  const World& W = *(World*)w;
  for (int i = 0; i < n; ++i) {
    RayHit h = world_ray_cast_closest(W, Vec2(p1p2[4 * i], p1p2[4 * i + 1]), Vec2(p1p2[4 * i + 2], p1p2[4 * i + 3]));
    std::memset(&out[i], 0, sizeof(out[i]));
    out[i].fixture = h.fixture;
    out[i].child_index = h.child;
    out[i].fraction = h.fraction;
    out[i].point_x = h.point.x; out[i].point_y = h.point.y;
    out[i].normal_x = h.normal.x; out[i].normal_y = h.normal.y;
  }
}
void b2o_query_aabb(void* w, const float* boxes, int n, int max_hits, int* counts, int* hits) {
  const World& W = *(World*)w;
  for (int i = 0; i < n; ++i) {
    AABB box;
    box.lower = Vec2(boxes[4 * i], boxes[4 * i + 1]);
    box.upper = Vec2(boxes[4 * i + 2], boxes[4 * i + 3]);
    std::vector<std::pair<int, int>> found;
    world_query_aabb(W, box, found);
    counts[i] = (int)found.size();
    for (int k = 0; k < (int)found.size() && k < max_hits; ++k) {
      hits[((size_t)i * max_hits + k) * 2] = found[k].first;
      hits[((size_t)i * max_hits + k) * 2 + 1] = found[k].second;
    }
  }
}

// ---- snapshot export in the b2gpu.h format
static int total_shapes(const World& W) {
  int n = 0;
  for (auto& f : W.fixtures) n += f.shape.child_count();
  return n;
}
void b2o_snapshot_sizes(void* w, b2gpu_snapshot_sizes* n) {
  const World& W = *(World*)w;
  n->body_count = (int)W.bodies.size();
  n->fixture_count = (int)W.fixtures.size();
  n->shape_count = total_shapes(W);
  n->proxy_count = (int)W.proxies.size();
  n->node_count = W.broad_phase.tree.node_capacity;
  n->contact_count = W.contact_count;
  n->move_count = (int)W.broad_phase.move_buffer.size();
  n->joint_count = (int)W.joints.size();
}
static void fill_shape_rec(const Shape& s, b2gpu_shape_rec* r) {
  std::memset(r, 0, sizeof(*r));
  r->type = s.type;
  r->radius = s.radius;
  if (s.type == E_CIRCLE) { r->cx = s.p.x; r->cy = s.p.y; r->v[0] = s.p.x; r->v[1] = s.p.y; }
  else if (s.type == E_EDGE) {
    r->one_sided = s.one_sided ? 1 : 0;
    r->v[0] = s.v0.x; r->v[1] = s.v0.y; r->v[2] = s.v1.x; r->v[3] = s.v1.y;
    r->v[4] = s.v2.x; r->v[5] = s.v2.y; r->v[6] = s.v3.x; r->v[7] = s.v3.y;
  } else {
    r->count = s.count;
    r->cx = s.centroid.x; r->cy = s.centroid.y;
    for (int i = 0; i < MAX_POLYGON_VERTICES; ++i) {
      r->v[2 * i] = s.vertices[i].x; r->v[2 * i + 1] = s.vertices[i].y;
      r->n[2 * i] = s.normals[i].x; r->n[2 * i + 1] = s.normals[i].y;
    }
  }
}
int b2o_snapshot_export(void* w, b2gpu_snapshot* out) {
  World& W = *(World*)w;
  b2gpu_snapshot_sizes need;
  b2o_snapshot_sizes(w, &need);
  if (out->n.body_count < need.body_count || out->n.fixture_count < need.fixture_count ||
      out->n.shape_count < need.shape_count || out->n.proxy_count < need.proxy_count ||
      out->n.node_count < need.node_count || out->n.contact_count < need.contact_count ||
      out->n.move_count < need.move_count || out->n.joint_count < need.joint_count || (need.joint_count > 0 && !out->joints))
    return -1;
  out->n = need;
  for (size_t i = 0; i < W.joints.size(); ++i) {
    const Joint& j = W.joints[i];
    b2gpu_joint_rec& r = out->joints[i];
    std::memset(&r, 0, sizeof(r));
    r.type = j.type; r.body_a = j.body_a; r.body_b = j.body_b;
    r.flags = (j.collide_connected ? B2GPU_JOINT_COLLIDE_CONNECTED : 0) | (j.enable_limit ? B2GPU_JOINT_ENABLE_LIMIT : 0) |
              (j.enable_motor ? B2GPU_JOINT_ENABLE_MOTOR : 0);
    r.local_anchor_a[0] = j.local_anchor_a.x; r.local_anchor_a[1] = j.local_anchor_a.y;
    r.local_anchor_b[0] = j.local_anchor_b.x; r.local_anchor_b[1] = j.local_anchor_b.y;
    if (j.type == J_REVOLUTE) {
      r.param[0] = j.reference_angle; r.param[1] = j.lower_angle; r.param[2] = j.upper_angle;
      r.param[3] = j.max_motor_torque; r.param[4] = j.motor_speed;
      r.impulse[0] = j.impulse2.x; r.impulse[1] = j.impulse2.y; r.impulse[2] = j.motor_impulse;
    } else if (j.type == J_PRISMATIC) {
      r.param[0] = j.reference_angle; r.param[1] = j.lower_angle; r.param[2] = j.upper_angle;
      r.param[3] = j.max_motor_torque; r.param[4] = j.motor_speed;
      r.param[5] = j.local_xaxis_a.x; r.param[6] = j.local_xaxis_a.y;
      r.impulse[0] = j.impulse2.x; r.impulse[1] = j.impulse2.y; r.impulse[2] = j.motor_impulse;
    } else if (j.type == J_FRICTION || j.type == J_MOTOR) {
      r.param[0] = j.max_force; r.param[1] = j.max_motor_torque;
      if (j.type == J_MOTOR) { r.param[2] = j.reference_angle; r.param[3] = j.correction_factor; }
      r.impulse[0] = j.impulse2.x; r.impulse[1] = j.impulse2.y; r.impulse[2] = j.motor_impulse;
    } else if (j.type == J_GEAR) {  // b2gpu.h: the static part overflows into impulse[1..6]
      r.flags |= (j.type_a == J_PRISMATIC ? B2GPU_JOINT_GEAR_PRISMATIC_1 : 0) | (j.type_b == J_PRISMATIC ? B2GPU_JOINT_GEAR_PRISMATIC_2 : 0);
      r.param[0] = j.local_anchor_c.x; r.param[1] = j.local_anchor_c.y; r.param[2] = j.local_anchor_d.x; r.param[3] = j.local_anchor_d.y;
      r.param[4] = j.local_axis_c.x; r.param[5] = j.local_axis_c.y; r.param[6] = j.local_axis_d.x; r.param[7] = j.local_axis_d.y;
      r.impulse[0] = j.impulse; r.impulse[1] = j.reference_angle; r.impulse[2] = j.reference_angle_b;
      r.impulse[3] = j.constant; r.impulse[4] = j.ratio;
      std::memcpy(&r.impulse[5], &j.body_c, 4); std::memcpy(&r.impulse[6], &j.body_d, 4);
    } else if (j.type == J_PULLEY) {
      r.param[0] = j.ground_anchor_a.x; r.param[1] = j.ground_anchor_a.y; r.param[2] = j.ground_anchor_b.x; r.param[3] = j.ground_anchor_b.y;
      r.param[4] = j.length; r.param[5] = j.length_b; r.param[6] = j.ratio; r.param[7] = j.constant;
      r.impulse[0] = j.impulse;
    } else if (j.type == J_MOUSE) {
      r.param[0] = j.max_force; r.param[1] = j.stiffness; r.param[2] = j.damping;
      r.param[3] = j.ground_anchor_a.x; r.param[4] = j.ground_anchor_a.y;  // the target: per world, like the motor settings
      r.impulse[0] = j.impulse2.x; r.impulse[1] = j.impulse2.y;
    } else if (j.type == J_WHEEL) {
      r.param[0] = j.stiffness; r.param[1] = j.lower_angle; r.param[2] = j.upper_angle;
      r.param[3] = j.max_motor_torque; r.param[4] = j.motor_speed;
      r.param[5] = j.local_xaxis_a.x; r.param[6] = j.local_xaxis_a.y; r.param[7] = j.damping;
      r.impulse[0] = j.impulse; r.impulse[1] = j.spring_impulse; r.impulse[2] = j.motor_impulse;
    } else if (j.type == J_WELD) {
      r.param[0] = j.reference_angle; r.param[3] = j.stiffness; r.param[4] = j.damping;
      r.impulse[0] = j.impulse3[0]; r.impulse[1] = j.impulse3[1]; r.impulse[2] = j.impulse3[2];
    } else {
      r.param[0] = j.length; r.param[1] = j.min_length; r.param[2] = j.max_length; r.param[3] = j.stiffness; r.param[4] = j.damping;
      r.impulse[0] = j.impulse;
    }
    if (j.type != J_GEAR) { r.impulse[3] = j.lower_impulse; r.impulse[4] = j.upper_impulse; }
  }
  b2gpu_world_rec& wr = out->world;
  std::memset(&wr, 0, sizeof(wr));
  wr.gravity_x = W.gravity.x; wr.gravity_y = W.gravity.y;
  wr.inv_dt0 = W.inv_dt0;
  wr.flags = (W.allow_sleep ? B2GPU_WORLD_ALLOW_SLEEP : 0) | (W.warm_starting ? B2GPU_WORLD_WARM_STARTING : 0) |
             (W.new_contacts ? B2GPU_WORLD_NEW_CONTACTS : 0) | (W.clear_forces_flag ? B2GPU_WORLD_CLEAR_FORCES : 0) |
             (W.block_solve ? B2GPU_WORLD_BLOCK_SOLVE : 0);
  const DynamicTree& T = W.broad_phase.tree;
  wr.tree_root = T.root; wr.tree_free_list = T.free_list; wr.tree_node_count = T.node_count;
  wr.tree_node_capacity = T.node_capacity; wr.tree_insertion_count = T.insertion_count;
  wr.proxy_count = W.broad_phase.proxy_count;
  for (size_t i = 0; i < W.bodies.size(); ++i) {
    const Body& b = W.bodies[i];
    b2gpu_body_rec& r = out->bodies[i];
    std::memset(&r, 0, sizeof(r));
    r.type = b.type; r.flags = b.flags;
    r.xf_px = b.xf.p.x; r.xf_py = b.xf.p.y; r.xf_qs = b.xf.q.s; r.xf_qc = b.xf.q.c;
    r.lc_x = b.sweep.local_center.x; r.lc_y = b.sweep.local_center.y;
    r.c0_x = b.sweep.c0.x; r.c0_y = b.sweep.c0.y; r.c_x = b.sweep.c.x; r.c_y = b.sweep.c.y;
    r.a0 = b.sweep.a0; r.a = b.sweep.a;
    r.vx = b.linear_velocity.x; r.vy = b.linear_velocity.y; r.w = b.angular_velocity;
    r.fx = b.force.x; r.fy = b.force.y; r.torque = b.torque;
    r.mass = b.mass; r.inv_mass = b.inv_mass; r.inertia = b.i; r.inv_inertia = b.inv_i;
    r.linear_damping = b.linear_damping; r.angular_damping = b.angular_damping; r.gravity_scale = b.gravity_scale;
    r.sleep_time = b.sleep_time;
    r.fixture_head = b.fixture_list; r.fixture_count = b.fixture_count;
  }
  int si = 0;
  for (size_t i = 0; i < W.fixtures.size(); ++i) {
    const Fixture& f = W.fixtures[i];
    b2gpu_fixture_rec& r = out->fixtures[i];
    std::memset(&r, 0, sizeof(r));
    r.body = f.body; r.next = f.next; r.shape_type = f.shape.type; r.shape_first = si;
    r.child_count = f.shape.child_count(); r.proxy_first = f.proxy_count > 0 ? f.proxy_first : -1;
    r.density = f.density; r.friction = f.friction; r.restitution = f.restitution;
    r.restitution_threshold = f.restitution_threshold;
    r.category_bits = f.filter.category_bits; r.mask_bits = f.filter.mask_bits; r.group_index = f.filter.group_index;
    r.is_sensor = f.is_sensor ? 1 : 0;
    if (f.shape.type == E_CHAIN) {
      for (int c = 0; c < r.child_count; ++c) {
        Shape e; chain_get_child_edge(f.shape, e, c);
        fill_shape_rec(e, &out->shapes[si++]);
      }
    } else {
      fill_shape_rec(f.shape, &out->shapes[si++]);
    }
  }
  for (size_t i = 0; i < W.proxies.size(); ++i) {
    const FixtureProxy& p = W.proxies[i];
    b2gpu_proxy_rec& r = out->proxies[i];
    r.fixture = p.fixture; r.child_index = p.child_index; r.proxy_id = p.proxy_id; r.reserved = 0;
    r.aabb[0] = p.aabb.lower.x; r.aabb[1] = p.aabb.lower.y; r.aabb[2] = p.aabb.upper.x; r.aabb[3] = p.aabb.upper.y;
  }
  for (int i = 0; i < T.node_capacity; ++i) {
    const TreeNode& nd = T.nodes[i];
    b2gpu_tree_node_rec& r = out->nodes[i];
    r.aabb[0] = nd.aabb.lower.x; r.aabb[1] = nd.aabb.lower.y; r.aabb[2] = nd.aabb.upper.x; r.aabb[3] = nd.aabb.upper.y;
    r.parent = nd.parent; r.child1 = nd.child1; r.child2 = nd.child2; r.height = nd.height;
    r.proxy = nd.user_data; r.moved = nd.moved ? 1 : 0;
  }
  {
    std::vector<int> order;
    for (int c = W.contact_list; c != -1; c = W.contacts[c].next) order.push_back(c);
    int k = 0;
    for (auto it = order.rbegin(); it != order.rend(); ++it, ++k) {
      const Contact& c = W.contacts[*it];
      b2gpu_contact_rec& r = out->contacts[k];
      std::memset(&r, 0, sizeof(r));
      r.fixture_a = c.fixture_a; r.fixture_b = c.fixture_b; r.index_a = c.index_a; r.index_b = c.index_b;
      r.flags = c.flags;
      r.friction = c.friction; r.restitution = c.restitution; r.restitution_threshold = c.restitution_threshold;
      r.tangent_speed = c.tangent_speed;
      const Manifold& m = c.manifold;
      for (int j = 0; j < 2; ++j) {
        r.manifold.points[j].lp_x = m.points[j].local_point.x; r.manifold.points[j].lp_y = m.points[j].local_point.y;
        r.manifold.points[j].normal_impulse = m.points[j].normal_impulse;
        r.manifold.points[j].tangent_impulse = m.points[j].tangent_impulse;
        r.manifold.points[j].id = m.points[j].id.key();
      }
      r.manifold.ln_x = m.local_normal.x; r.manifold.ln_y = m.local_normal.y;
      r.manifold.lp_x = m.local_point.x; r.manifold.lp_y = m.local_point.y;
      r.manifold.type = m.type; r.manifold.point_count = m.point_count;
    }
  }
  for (size_t i = 0; i < W.broad_phase.move_buffer.size(); ++i) out->move_buffer[i] = W.broad_phase.move_buffer[i];
  return 0;
}

// Compact per-body state [body][8] = c.x c.y a v.x v.y w xf.p.x xf.p.y (b2gpu_batch_get_body_state layout)
void b2o_get_body_state(void* w, float* out) {
  const World& W = *(World*)w;
  for (size_t i = 0; i < W.bodies.size(); ++i) {
    const Body& b = W.bodies[i];
    float* o = out + 8 * i;
    o[0] = b.sweep.c.x; o[1] = b.sweep.c.y; o[2] = b.sweep.a; o[3] = b.linear_velocity.x; o[4] = b.linear_velocity.y;
    o[5] = b.angular_velocity; o[6] = b.xf.p.x; o[7] = b.xf.p.y;
  }
}

// ---- CPU baseline: one world per host thread. Returns seconds of wall time.
// A persistent pool (created on first use, reused by every call: the timed region contains no thread creation), worlds
// dealt out statically round-robin so every thread steps the same number of worlds when n_worlds % threads == 0.
namespace {
struct Pool {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  long generation = 0;
  int pending = 0;
  bool quit = false;
  void** worlds = nullptr;
  int n_worlds = 0, steps = 0, vi = 0, pi = 0;
  float dt = 0.0f;
  void worker(int t, int nt, long seen) {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m);
        cv_go.wait(lk, [&] { return quit || generation != seen; });
        if (quit) return;
        seen = generation;
      }
      for (int i = t; i < n_worlds; i += nt) {
        World* W = (World*)worlds[i];
        for (int s = 0; s < steps; ++s) W->step(dt, vi, pi);
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
  void ensure(int nt) {
    if ((int)threads.size() == nt) return;
    shutdown();
    quit = false;
    const long gen = generation;
    for (int t = 0; t < nt; ++t) threads.emplace_back([this, t, nt, gen] { worker(t, nt, gen); });
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
    }
    cv_go.notify_all();
    for (auto& th : threads) th.join();
    threads.clear();
  }
  ~Pool() { shutdown(); }
};
Pool g_pool;
}  // namespace
double b2o_run_worlds_mt(void** worlds, int n_worlds, int steps, float dt, int vi, int pi, int threads) {
  if (threads < 1) threads = 1;
  g_pool.ensure(threads);
  double t0 = now_ms();
  {
    std::lock_guard<std::mutex> lk(g_pool.m);
    g_pool.worlds = worlds; g_pool.n_worlds = n_worlds; g_pool.steps = steps; g_pool.dt = dt; g_pool.vi = vi; g_pool.pi = pi;
    g_pool.pending = threads;
    ++g_pool.generation;
  }
  g_pool.cv_go.notify_all();
  {
    std::unique_lock<std::mutex> lk(g_pool.m);
    g_pool.cv_done.wait(lk, [&] { return g_pool.pending == 0; });
  }
  return (now_ms() - t0) * 1e-3;
}
// libm sinf/cosf of an array (what f32::sin / f32::cos lower to on linux-gnu): reference for the
// device trigonometry tests.
void b2o_sincosf(const float* in, float* s, float* c, int n) {
  for (int i = 0; i < n; ++i) { s[i] = sinf(in[i]); c[i] = cosf(in[i]); }
}
int b2o_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
