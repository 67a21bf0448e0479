// TEST INFRASTRUCTURE — CPU oracle (see b2o_math.hpp header). PARITY UNPINNED beyond the
// reference's own tests (tests/test.rs, world_test.rs, collision_test.rs, math_test.rs).
//
// b2o_world.hpp — restates the step path of box2d-rs with index-based storage:
//   src/private/dynamics/b2_world.rs (step :903-959, solve :356-531, create_body :83-98)
//   src/private/dynamics/b2_body.rs, src/b2_body.rs (set_awake, synchronize_*)
//   src/private/dynamics/b2_fixture.rs (create_proxies, synchronize)
//   src/private/dynamics/b2_contact_manager.rs, b2_contact.rs, b2_contact_registers.rs
//   src/private/dynamics/b2_island_private.rs, b2_contact_solver_private.rs
// Intrusive lists are kept as real push_front linked lists through indices so the
// iteration orders of SURVEY.md §3.4 hold by construction.
#pragma once
#include <algorithm>
#include <chrono>
#include <vector>

#include "b2o_collision.hpp"
#include "b2o_distance.hpp"
#include "b2o_tree.hpp"

namespace b2o {

enum BodyType { STATIC_BODY = 0, KINEMATIC_BODY = 1, DYNAMIC_BODY = 2 };
enum : uint32_t {  // src/b2_body.rs:209-221
  BF_ISLAND = 0x0001, BF_AWAKE = 0x0002, BF_AUTO_SLEEP = 0x0004, BF_BULLET = 0x0008,
  BF_FIXED_ROTATION = 0x0010, BF_ENABLED = 0x0020, BF_TOI = 0x0040
};
enum : uint32_t {  // src/b2_contact.rs:283-305
  CF_ISLAND = 0x0001, CF_TOUCHING = 0x0002, CF_ENABLED = 0x0004, CF_FILTER = 0x0008
};

struct BodyDef {  // src/b2_body.rs:39-58
  int type = STATIC_BODY;
  Vec2 position;
  float angle = 0.0f;
  Vec2 linear_velocity;
  float angular_velocity = 0.0f, linear_damping = 0.0f, angular_damping = 0.0f;
  bool allow_sleep = true, awake = true, fixed_rotation = false, bullet = false, enabled = true;
  float gravity_scale = 1.0f;
};
struct Filter {
  uint16_t category_bits = 0x0001, mask_bits = 0xFFFF;
  int16_t group_index = 0;
};
struct FixtureDef {  // src/b2_fixture.rs:44-57
  float friction = 0.2f, restitution = 0.0f, restitution_threshold = 1.0f * LENGTH_UNITS_PER_METER, density = 0.0f;
  bool is_sensor = false;
  Filter filter;
};

struct Body {
  int type = STATIC_BODY;
  uint32_t flags = 0;
  int island_index = -1;
  Transform xf;
  Sweep sweep;
  Vec2 linear_velocity;
  float angular_velocity = 0.0f;
  Vec2 force;
  float torque = 0.0f;
  int prev = -1, next = -1;        // world body list
  int fixture_list = -1;            // head = newest
  int fixture_count = 0;
  int contact_list = -1;            // head edge id (2*contact + side)
  std::vector<int> joint_edges;     // m_joint_list in push order (2*joint + side); the list iterates it in reverse
  float mass = 0.0f, inv_mass = 0.0f, i = 0.0f, inv_i = 0.0f;
  float linear_damping = 0.0f, angular_damping = 0.0f, gravity_scale = 1.0f, sleep_time = 0.0f;
};
struct FixtureProxy {
  AABB aabb;
  int fixture = -1, child_index = 0, proxy_id = -1;
};
struct Fixture {
  int body = -1, next = -1;
  Shape shape;
  float density = 0.0f, friction = 0.0f, restitution = 0.0f, restitution_threshold = 0.0f;
  Filter filter;
  bool is_sensor = false;
  int proxy_first = -1, proxy_count = 0;  // into World::proxies
};
struct ContactEdge {
  int other = -1, prev = -1, next = -1;
};
struct Contact {
  bool alive = false;
  uint32_t flags = 0;
  int prev = -1, next = -1;  // world contact list
  ContactEdge node_a, node_b;
  int fixture_a = -1, fixture_b = -1, index_a = 0, index_b = 0;
  Manifold manifold;
  float friction = 0.0f, restitution = 0.0f, restitution_threshold = 0.0f, tangent_speed = 0.0f;
};

// Joints: B2jointDef + B2revoluteJointDef / B2distanceJointDef as one plain struct
// (src/b2_joint.rs:112-122, src/joints/b2_revolute_joint.rs:10-72, src/joints/b2_distance_joint.rs:11-58).
enum JointType { J_DISTANCE = 1, J_FRICTION = 2, J_GEAR = 3, J_MOTOR = 4, J_MOUSE = 5, J_PRISMATIC = 6, J_PULLEY = 7, J_REVOLUTE = 8, J_WELD = 9, J_WHEEL = 10 };  // B2jointType numbering (src/b2_joint.rs:46-58)
struct JointDef {
  int type = 0, body_a = -1, body_b = -1;
  bool collide_connected = false;
  Vec2 local_anchor_a, local_anchor_b;
  float reference_angle = 0.0f, lower_angle = 0.0f, upper_angle = 0.0f, max_motor_torque = 0.0f, motor_speed = 0.0f;
  bool enable_limit = false, enable_motor = false;
  float length = 1.0f, min_length = 0.0f, max_length = MAX_FLOAT, stiffness = 0.0f, damping = 0.0f;
  // prismatic (src/joints/b2_prismatic_joint.rs:10-72): lower_angle / upper_angle carry the translation limits,
  // max_motor_torque the maximum motor force
  Vec2 local_axis_a = Vec2(1.0f, 0.0f);
  // friction (src/joints/b2_friction_joint.rs:9-50) / motor (src/joints/b2_motor_joint.rs:9-59): max_force, and for the motor
  // joint linear_offset (in local_anchor_a), angular_offset (in reference_angle), correction_factor; max_motor_torque = max_torque
  float max_force = 0.0f, correction_factor = 0.3f;
  // pulley (src/joints/b2_pulley_joint.rs:10-78): ground anchors, rest lengths (length = length_a), ratio
  // mouse (src/joints/b2_mouse_joint.rs:8-50): target (world point); max_force, stiffness, damping as named
  Vec2 ground_anchor_a = Vec2(-1.0f, 1.0f), ground_anchor_b = Vec2(1.0f, 1.0f), target;
  float length_b = 0.0f, ratio = 1.0f;
  // gear (src/joints/b2_gear_joint.rs:12-40): the two revolute / prismatic joints it couples (indices), `ratio`
  int joint1 = -1, joint2 = -1;
};
struct Joint {  // B2joint + B2revoluteJoint (src/joints/b2_revolute_joint.rs:104-136) / B2distanceJoint fields
  int type = 0, body_a = -1, body_b = -1;
  bool collide_connected = false, island_flag = false;
  Vec2 local_anchor_a, local_anchor_b;
  // revolute: solver shared
  Vec2 impulse2;
  float motor_impulse = 0.0f, lower_impulse = 0.0f, upper_impulse = 0.0f;
  bool enable_motor = false, enable_limit = false;
  float max_motor_torque = 0.0f, motor_speed = 0.0f, reference_angle = 0.0f, lower_angle = 0.0f, upper_angle = 0.0f;
  // distance: solver shared
  float length = 0.0f, min_length = 0.0f, max_length = 0.0f, stiffness = 0.0f, damping = 0.0f, impulse = 0.0f;
  float gamma = 0.0f, bias = 0.0f, current_length = 0.0f, mass = 0.0f, soft_mass = 0.0f;
  Vec2 u;
  // prismatic (src/joints/b2_prismatic_joint.rs:101-136): shares impulse2 / motor / limit impulses and switches with the
  // revolute joint (lower_angle / upper_angle = translation limits, max_motor_torque = max motor force)
  Vec2 local_xaxis_a, local_yaxis_a, axis, perp;
  float s1 = 0.0f, s2 = 0.0f, a1 = 0.0f, a2 = 0.0f, translation = 0.0f;
  // wheel (src/joints/b2_wheel_joint.rs:120-170): impulse (scalar, here `impulse`), spring_impulse; shares the motor / limit
  // impulses and switches, local_xaxis_a / local_yaxis_a, translation, gamma, bias, mass, axial_mass
  float spring_impulse = 0.0f, spring_mass = 0.0f, motor_mass = 0.0f, s_ax = 0.0f, s_bx = 0.0f, s_ay = 0.0f, s_by = 0.0f;
  Vec2 ax, ay;
  // friction / motor: impulse2 = linear impulse, motor_impulse = angular impulse, max_motor_torque = max torque;
  // k = linear mass (inverse), axial_mass = angular mass
  float max_force = 0.0f, correction_factor = 0.0f, angular_error = 0.0f;
  Vec2 linear_error;
  // pulley (private joints/b2_pulley_joint.rs:8-44): length = length_a; `constant` = length_a + ratio * length_b; u = u_a
  // mouse (src/joints/b2_mouse_joint.rs:140-170): ground_anchor_a = target, impulse2, gamma, `beta`, linear_error = C, k = mass
  Vec2 ground_anchor_a, ground_anchor_b, u_b;
  float length_b = 0.0f, ratio = 1.0f, constant = 0.0f, beta = 0.0f;
  // gear (src/joints/b2_gear_joint.rs:150-200): bodies C, D = body A of joint 1 / 2 (A, B = their body B), the joints' types,
  // anchors and axes copied at creation; reference_angle = reference_angle_a; solver temp jv_ac, jv_bd, jw_a..d, `mass`
  int body_c = -1, body_d = -1, type_a = 0, type_b = 0;
  Vec2 local_anchor_c, local_anchor_d, local_axis_c, local_axis_d, jv_ac, jv_bd;
  float reference_angle_b = 0.0f, jw_a = 0.0f, jw_b = 0.0f, jw_c = 0.0f, jw_d = 0.0f;
  // weld (src/joints/b2_weld_joint.rs:66-90): impulse (x, y, angular), effective mass B2Mat33 as ex.xyz ey.xyz ez.xyz
  float impulse3[3] = {0.0f, 0.0f, 0.0f}, m33[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  // solver temp
  int index_a = 0, index_b = 0;
  Vec2 r_a, r_b, local_center_a, local_center_b;
  float inv_mass_a = 0.0f, inv_mass_b = 0.0f, inv_ia = 0.0f, inv_ib = 0.0f;
  Mat22 k;
  float angle = 0.0f, axial_mass = 0.0f;
};

struct TimeStep {  // src/b2_time_step.rs
  float dt = 0.0f, inv_dt = 0.0f, dt_ratio = 0.0f;
  int velocity_iterations = 0, position_iterations = 0;
  bool warm_starting = false;
};
struct Profile {  // src/b2_time_step.rs:5-15 (ms)
  double step = 0, collide = 0, solve = 0, solve_init = 0, solve_velocity = 0, solve_position = 0, broadphase = 0;
};
// B2contactListener::begin_contact / end_contact (src/b2_world_callbacks.rs:68-104) as the reference would fire them
// during one step, recorded in firing order (test infrastructure for b2gpu_contact_events).
struct ContactEvent {
  int type;  // 1 = begin_contact, 2 = end_contact
  int fixture_a, index_a, fixture_b, index_b;
};
// B2contactListener::post_solve (src/b2_world_callbacks.rs:94-103) as B2island::report would call it
// (b2_island_private.rs:460-487): per island, contacts in island order, the impulses of the velocity constraint.
struct PostSolveEvent {
  int fixture_a, index_a, fixture_b, index_b, count;
  float normal_impulses[2], tangent_impulses[2];
};
struct StepStats {
  int contacts = 0, touching = 0, destroyed = 0, islands = 0, island_bodies = 0, island_contacts = 0, moved = 0, pairs = 0,
      created = 0, awake_bodies = 0, solver_levels = 0;
};

// src/private/dynamics/b2_contact_solver.rs:11-119
struct VelocityConstraintPoint {
  Vec2 r_a, r_b;
  float normal_impulse = 0, tangent_impulse = 0, normal_mass = 0, tangent_mass = 0, velocity_bias = 0;
};
struct ContactVelocityConstraint {
  VelocityConstraintPoint points[MAX_MANIFOLD_POINTS];
  Vec2 normal;
  Mat22 normal_mass, k;
  int index_a = 0, index_b = 0;
  float inv_mass_a = 0, inv_mass_b = 0, inv_ia = 0, inv_ib = 0, friction = 0, restitution = 0, threshold = 0, tangent_speed = 0;
  int point_count = 0, contact_index = 0;
};
struct ContactPositionConstraint {
  Vec2 local_points[MAX_MANIFOLD_POINTS];
  Vec2 local_normal, local_point;
  int index_a = 0, index_b = 0;
  float inv_mass_a = 0, inv_mass_b = 0;
  Vec2 local_center_a, local_center_b;
  float inv_ia = 0, inv_ib = 0;
  int type = 0;
  float radius_a = 0, radius_b = 0;
  int point_count = 0;
};
struct Position { Vec2 c; float a = 0; };
struct Velocity { Vec2 v; float w = 0; };

inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct World {
  // b2_world.rs(private):26-56
  Vec2 gravity;
  bool warm_starting = true, continuous_physics = false /* TOI out of scope */, allow_sleep = true;
  bool new_contacts = false, locked = false, clear_forces_flag = true, step_complete = true;
  bool block_solve = true;  // G_BLOCK_SOLVE
  float inv_dt0 = 0.0f;
  std::vector<Body> bodies;        // index = creation order (no destroy_body in scope)
  int body_list = -1;              // head = newest
  std::vector<Fixture> fixtures;   // creation order
  std::vector<FixtureProxy> proxies;
  std::vector<Contact> contacts;   // slots
  std::vector<int> contact_free;
  int contact_list = -1, contact_count = 0;
  BroadPhase broad_phase;
  Profile profile;
  StepStats stats;
  bool collect_levels = false;     // compute the wavefront depth statistic (diagnostic only)
  bool collect_dag = false;        // also the largest island's DAG statistics (dag_collect)
  double dag_handover = 2.0;

  explicit World(Vec2 g) : gravity(g) {}

  // ---------------------------------------------------------------- bodies
  int create_body(const BodyDef& bd) {  // b2_world.rs(private):83-98 ; b2_body.rs(private):15-90
    Body b;
    if (bd.bullet) b.flags |= BF_BULLET;
    if (bd.fixed_rotation) b.flags |= BF_FIXED_ROTATION;
    if (bd.allow_sleep) b.flags |= BF_AUTO_SLEEP;
    if (bd.awake && bd.type != STATIC_BODY) b.flags |= BF_AWAKE;
    if (bd.enabled) b.flags |= BF_ENABLED;
    b.xf.p = bd.position;
    b.xf.q = Rot(bd.angle);
    b.sweep.local_center.set_zero();
    b.sweep.c0 = b.xf.p;
    b.sweep.c = b.xf.p;
    b.sweep.a0 = bd.angle;
    b.sweep.a = bd.angle;
    b.linear_velocity = bd.linear_velocity;
    b.angular_velocity = bd.angular_velocity;
    b.linear_damping = bd.linear_damping;
    b.angular_damping = bd.angular_damping;
    b.gravity_scale = bd.gravity_scale;
    b.type = bd.type;
    int id = (int)bodies.size();
    b.next = body_list;  // push_front
    if (body_list != -1) bodies[body_list].prev = id;
    bodies.push_back(b);
    body_list = id;
    return id;
  }
  void set_awake(int bi, bool flag) {  // src/b2_body.rs:783-801
    Body& b = bodies[bi];
    if (b.type == STATIC_BODY) return;
    if (flag) {
      b.flags |= BF_AWAKE;
      b.sleep_time = 0.0f;
    } else {
      b.flags &= ~BF_AWAKE;
      b.sleep_time = 0.0f;
      b.linear_velocity.set_zero();
      b.angular_velocity = 0.0f;
      b.force.set_zero();
      b.torque = 0.0f;
    }
  }
  void synchronize_transform(Body& b) {  // src/b2_body.rs:974-977
    b.xf.q.set(b.sweep.a);
    b.xf.p = b.sweep.c - b2_mul_rot(b.xf.q, b.sweep.local_center);
  }
  void reset_mass_data(int bi) {  // b2_body.rs(private):292-350
    Body& b = bodies[bi];
    b.mass = 0.0f; b.inv_mass = 0.0f; b.i = 0.0f; b.inv_i = 0.0f;
    b.sweep.local_center.set_zero();
    if (b.type == STATIC_BODY || b.type == KINEMATIC_BODY) {
      b.sweep.c0 = b.xf.p;
      b.sweep.c = b.xf.p;
      b.sweep.a0 = b.sweep.a;
      return;
    }
    Vec2 local_center(0.0f, 0.0f);
    for (int f = b.fixture_list; f != -1; f = fixtures[f].next) {
      const Fixture& fx = fixtures[f];
      if (fx.density == 0.0f) continue;
      MassData md;
      shape_compute_mass(fx.shape, md, fx.density);
      b.mass += md.mass;
      local_center += md.mass * md.center;
      b.i += md.i;
    }
    if (b.mass > 0.0f) {
      b.inv_mass = 1.0f / b.mass;
      local_center *= b.inv_mass;
    }
    if (b.i > 0.0f && !(b.flags & BF_FIXED_ROTATION)) {
      b.i -= b.mass * b2_dot(local_center, local_center);
      b.inv_i = 1.0f / b.i;
    } else {
      b.i = 0.0f;
      b.inv_i = 0.0f;
    }
    Vec2 old_center = b.sweep.c;
    b.sweep.local_center = local_center;
    b.sweep.c0 = b2_mul_xf(b.xf, b.sweep.local_center);
    b.sweep.c = b.sweep.c0;
    b.linear_velocity += b2_cross_sv(b.angular_velocity, b.sweep.c - old_center);
  }
  int create_fixture(int bi, const FixtureDef& def, const Shape& shape) {  // b2_body.rs(private):155-200
    Fixture fx;
    fx.friction = def.friction;
    fx.restitution = def.restitution;
    fx.restitution_threshold = def.restitution_threshold;
    fx.body = bi;
    fx.filter = def.filter;
    fx.is_sensor = def.is_sensor;
    fx.shape = shape;
    fx.density = def.density;
    int fi = (int)fixtures.size();
    fixtures.push_back(fx);
    Body& b = bodies[bi];
    if (b.flags & BF_ENABLED) create_proxies(fi, b.xf);
    fixtures[fi].next = b.fixture_list;  // push_front
    b.fixture_list = fi;
    b.fixture_count += 1;
    if (fixtures[fi].density > 0.0f) reset_mass_data(bi);
    new_contacts = true;
    return fi;
  }
  void create_proxies(int fi, const Transform& xf) {  // b2_fixture.rs(private):114-132
    Fixture& fx = fixtures[fi];
    fx.proxy_count = fx.shape.child_count();
    fx.proxy_first = (int)proxies.size();
    for (int i = 0; i < fx.proxy_count; ++i) {
      FixtureProxy p;
      shape_compute_aabb(fx.shape, p.aabb, xf, i);
      int pi = (int)proxies.size();
      p.proxy_id = broad_phase.create_proxy(p.aabb, pi);
      p.fixture = fi;
      p.child_index = i;
      proxies.push_back(p);
    }
  }
  void fixture_synchronize(int fi, const Transform& xf1, const Transform& xf2) {  // b2_fixture.rs(private):147-173
    Fixture& fx = fixtures[fi];
    if (fx.proxy_count == 0) return;
    for (int i = 0; i < fx.proxy_count; ++i) {
      FixtureProxy& p = proxies[fx.proxy_first + i];
      AABB aabb1, aabb2;
      shape_compute_aabb(fx.shape, aabb1, xf1, p.child_index);
      shape_compute_aabb(fx.shape, aabb2, xf2, p.child_index);
      p.aabb.combine_two(aabb1, aabb2);
      Vec2 displacement = aabb2.get_center() - aabb1.get_center();
      broad_phase.move_proxy(p.proxy_id, p.aabb, displacement);
    }
  }
  void synchronize_fixtures(int bi) {  // b2_body.rs(private):455-475
    Body& b = bodies[bi];
    if (b.flags & BF_AWAKE) {
      Transform xf1;
      xf1.q.set(b.sweep.a0);
      xf1.p = b.sweep.c0 - b2_mul_rot(xf1.q, b.sweep.local_center);
      for (int f = b.fixture_list; f != -1; f = fixtures[f].next) fixture_synchronize(f, xf1, b.xf);
    } else {
      for (int f = b.fixture_list; f != -1; f = fixtures[f].next) fixture_synchronize(f, b.xf, b.xf);
    }
  }
  void set_transform(int bi, Vec2 position, float angle) {  // b2_body.rs(private):418-444
    Body& b = bodies[bi];
    b.xf.q.set(angle);
    b.xf.p = position;
    b.sweep.c = b2_mul_xf(b.xf, b.sweep.local_center);
    b.sweep.a = angle;
    b.sweep.c0 = b.sweep.c;
    b.sweep.a0 = angle;
    for (int f = b.fixture_list; f != -1; f = fixtures[f].next) fixture_synchronize(f, b.xf, b.xf);
    new_contacts = true;
  }
  void set_linear_velocity(int bi, Vec2 v) {  // src/b2_body.rs set_linear_velocity
    Body& b = bodies[bi];
    if (b.type == STATIC_BODY) return;
    if (b2_dot(v, v) > 0.0f) set_awake(bi, true);
    b.linear_velocity = v;
  }
  void set_angular_velocity(int bi, float w) {
    Body& b = bodies[bi];
    if (b.type == STATIC_BODY) return;
    if (w * w > 0.0f) set_awake(bi, true);
    b.angular_velocity = w;
  }
  void apply_force_to_center(int bi, Vec2 f, bool wake) {  // src/b2_body.rs apply_force_to_center
    Body& b = bodies[bi];
    if (b.type != DYNAMIC_BODY) return;
    if (wake && !(b.flags & BF_AWAKE)) set_awake(bi, true);
    if (b.flags & BF_AWAKE) b.force += f;
  }
  // src/b2_body.rs:869-972 — the rest of the force / impulse API (inline module of B2body)
  void apply_force(int bi, Vec2 f, Vec2 point, bool wake) {  // :869-888
    Body& b = bodies[bi];
    if (b.type != DYNAMIC_BODY) return;
    if (wake && !(b.flags & BF_AWAKE)) set_awake(bi, true);
    if (b.flags & BF_AWAKE) {
      b.force += f;
      b.torque += b2_cross(point - b.sweep.c, f);
    }
  }
  void apply_torque(int bi, float torque, bool wake) {  // :905-918
    Body& b = bodies[bi];
    if (b.type != DYNAMIC_BODY) return;
    if (wake && !(b.flags & BF_AWAKE)) set_awake(bi, true);
    if (b.flags & BF_AWAKE) b.torque += torque;
  }
  void apply_linear_impulse(int bi, Vec2 impulse, Vec2 point, bool wake) {  // :920-939
    Body& b = bodies[bi];
    if (b.type != DYNAMIC_BODY) return;
    if (wake && !(b.flags & BF_AWAKE)) set_awake(bi, true);
    if (b.flags & BF_AWAKE) {
      b.linear_velocity += b.inv_mass * impulse;
      b.angular_velocity += b.inv_i * b2_cross(point - b.sweep.c, impulse);
    }
  }
  void apply_linear_impulse_to_center(int bi, Vec2 impulse, bool wake) {  // :941-957
    Body& b = bodies[bi];
    if (b.type != DYNAMIC_BODY) return;
    if (wake && !(b.flags & BF_AWAKE)) set_awake(bi, true);
    if (b.flags & BF_AWAKE) b.linear_velocity += b.inv_mass * impulse;
  }
  void apply_angular_impulse(int bi, float impulse, bool wake) {  // :959-972
    Body& b = bodies[bi];
    if (b.type != DYNAMIC_BODY) return;
    if (wake && !(b.flags & BF_AWAKE)) set_awake(bi, true);
    if (b.flags & BF_AWAKE) b.angular_velocity += b.inv_i * impulse;
  }
  bool body_should_collide(int self_, int other) const {  // b2_body.rs(private):391-416
    if (bodies[self_].type != DYNAMIC_BODY && bodies[other].type != DYNAMIC_BODY) return false;
    const std::vector<int>& je = bodies[self_].joint_edges;  // does a joint prevent collision?
    for (size_t i = je.size(); i-- > 0;) {
      const Joint& j = joints[je[i] >> 1];
      const int jn_other = (je[i] & 1) ? j.body_a : j.body_b;
      if (jn_other == other && !j.collide_connected) return false;
    }
    return true;
  }

  // ---------------------------------------------------------------- joints
  std::vector<Joint> joints;  // creation order
  // B2revoluteJointDef::initialize (src/joints/b2_revolute_joint.rs:77-85)
  JointDef revolute_joint_def(int body_a, int body_b, Vec2 anchor) const {
    JointDef d;
    d.type = J_REVOLUTE;
    d.body_a = body_a; d.body_b = body_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, anchor);  // get_local_point (src/b2_body.rs:728-730)
    d.local_anchor_b = b2_mul_t_xf(bodies[body_b].xf, anchor);
    d.reference_angle = bodies[body_b].sweep.a - bodies[body_a].sweep.a;
    return d;
  }
  // b2_distance_joint_def_initialize (private b2_distance_joint.rs:26-41)
  JointDef distance_joint_def(int b1, int b2, Vec2 anchor1, Vec2 anchor2) const {
    JointDef d;
    d.type = J_DISTANCE;
    d.body_a = b1; d.body_b = b2;
    d.local_anchor_a = b2_mul_t_xf(bodies[b1].xf, anchor1);
    d.local_anchor_b = b2_mul_t_xf(bodies[b2].xf, anchor2);
    Vec2 dd = anchor2 - anchor1;
    d.length = b2_max(dd.length(), LINEAR_SLOP);
    d.min_length = d.length;
    d.max_length = d.length;
    return d;
  }
  // B2prismaticJointDef::default + ::initialize (src/joints/b2_prismatic_joint.rs:10-89)
  JointDef prismatic_joint_def(int body_a, int body_b, Vec2 anchor, Vec2 axis) const {
    JointDef d;
    d.type = J_PRISMATIC;
    d.body_a = body_a; d.body_b = body_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, anchor);
    d.local_anchor_b = b2_mul_t_xf(bodies[body_b].xf, anchor);
    d.local_axis_a = b2_mul_t_rot(bodies[body_a].xf.q, axis);  // get_local_vector (src/b2_body.rs:733-735)
    d.reference_angle = bodies[body_b].sweep.a - bodies[body_a].sweep.a;
    return d;
  }
  // B2frictionJointDef::default + ::initialize (src/joints/b2_friction_joint.rs:9-50)
  JointDef friction_joint_def(int body_a, int body_b, Vec2 anchor) const {
    JointDef d;
    d.type = J_FRICTION;
    d.body_a = body_a; d.body_b = body_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, anchor);
    d.local_anchor_b = b2_mul_t_xf(bodies[body_b].xf, anchor);
    d.max_force = 0.0f; d.max_motor_torque = 0.0f;
    return d;
  }
  // B2pulleyJointDef::default + ::initialize (src/joints/b2_pulley_joint.rs:10-78): collide_connected defaults to true
  JointDef pulley_joint_def(int body_a, int body_b, Vec2 ground_a, Vec2 ground_b, Vec2 anchor_a, Vec2 anchor_b, float r) const {
    JointDef d;
    d.type = J_PULLEY;
    d.collide_connected = true;
    d.body_a = body_a; d.body_b = body_b;
    d.ground_anchor_a = ground_a; d.ground_anchor_b = ground_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, anchor_a);
    d.local_anchor_b = b2_mul_t_xf(bodies[body_b].xf, anchor_b);
    d.length = (anchor_a - ground_a).length();
    d.length_b = (anchor_b - ground_b).length();
    d.ratio = r;
    assert(r > EPSILON);
    return d;
  }
  // B2gearJointDef::default (src/joints/b2_gear_joint.rs:12-40) with joint1, joint2, ratio; bodies A / B as B2gearJoint::new
  // takes them (body B of joint 1 / joint 2)
  JointDef gear_joint_def(int joint1, int joint2, float r) const {
    JointDef d;
    d.type = J_GEAR;
    d.joint1 = joint1; d.joint2 = joint2;
    d.body_a = joints[joint1].body_b; d.body_b = joints[joint2].body_b;
    d.ratio = r;
    return d;
  }
  // B2mouseJointDef::default (src/joints/b2_mouse_joint.rs:8-21): target, max_force, stiffness, damping all zero
  JointDef mouse_joint_def(int body_a, int body_b, Vec2 target) const {
    JointDef d;
    d.type = J_MOUSE;
    d.body_a = body_a; d.body_b = body_b;
    d.target = target;
    d.length = 0.0f; d.min_length = 0.0f; d.max_length = 0.0f;
    return d;
  }
  // B2motorJointDef::default + ::initialize (src/joints/b2_motor_joint.rs:9-59)
  JointDef motor_joint_def(int body_a, int body_b) const {
    JointDef d;
    d.type = J_MOTOR;
    d.body_a = body_a; d.body_b = body_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, bodies[body_b].xf.p);  // linear_offset = get_local_point(body_b position)
    d.reference_angle = bodies[body_b].sweep.a - bodies[body_a].sweep.a;     // angular_offset
    d.max_force = 1.0f; d.max_motor_torque = 1.0f; d.correction_factor = 0.3f;
    return d;
  }
  // B2wheelJointDef::default + ::initialize (src/joints/b2_wheel_joint.rs:10-90)
  JointDef wheel_joint_def(int body_a, int body_b, Vec2 anchor, Vec2 axis) const {
    JointDef d;
    d.type = J_WHEEL;
    d.body_a = body_a; d.body_b = body_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, anchor);
    d.local_anchor_b = b2_mul_t_xf(bodies[body_b].xf, anchor);
    d.local_axis_a = b2_mul_t_rot(bodies[body_a].xf.q, axis);
    return d;
  }
  // B2weldJointDef::default + ::initialize (src/joints/b2_weld_joint.rs:10-50)
  JointDef weld_joint_def(int body_a, int body_b, Vec2 anchor) const {
    JointDef d;
    d.type = J_WELD;
    d.body_a = body_a; d.body_b = body_b;
    d.local_anchor_a = b2_mul_t_xf(bodies[body_a].xf, anchor);
    d.local_anchor_b = b2_mul_t_xf(bodies[body_b].xf, anchor);
    d.reference_angle = bodies[body_b].sweep.a - bodies[body_a].sweep.a;
    return d;
  }
  // b2_angular_stiffness (src/private/dynamics/b2_joint.rs:47-70); B2body::get_inertia (src/b2_body.rs:708-711)
  void angular_stiffness(float& stiffness, float& damping, float frequency_hertz, float damping_ratio, int body_a, int body_b) const {
    const Body& a = bodies[body_a];
    const Body& b = bodies[body_b];
    float ia = a.i + a.mass * b2_dot(a.sweep.local_center, a.sweep.local_center);
    float ib = b.i + b.mass * b2_dot(b.sweep.local_center, b.sweep.local_center);
    float i;
    if (ia > 0.0f && ib > 0.0f) i = ia * ib / (ia + ib);
    else if (ia > 0.0f) i = ia;
    else i = ib;
    float omega = 2.0f * PI * frequency_hertz;
    stiffness = i * omega * omega;
    damping = 2.0f * i * damping_ratio * omega;
  }
  // b2_linear_stiffness (src/private/dynamics/b2_joint.rs:22-45)
  void linear_stiffness(float& stiffness, float& damping, float frequency_hertz, float damping_ratio, int body_a, int body_b) const {
    float mass_a = bodies[body_a].mass, mass_b = bodies[body_b].mass, mass;
    if (mass_a > 0.0f && mass_b > 0.0f) mass = mass_a * mass_b / (mass_a + mass_b);
    else if (mass_a > 0.0f) mass = mass_a;
    else mass = mass_b;
    float omega = 2.0f * PI * frequency_hertz;
    stiffness = mass * omega * omega;
    damping = 2.0f * mass * damping_ratio * omega;
  }
  int create_joint(const JointDef& def) {  // b2_world.rs(private):156-262
    assert(def.body_a != def.body_b);
    Joint j;
    j.type = def.type; j.body_a = def.body_a; j.body_b = def.body_b; j.collide_connected = def.collide_connected;
    j.local_anchor_a = def.local_anchor_a; j.local_anchor_b = def.local_anchor_b;
    if (def.type == J_REVOLUTE) {  // B2revoluteJoint::new (src/joints/b2_revolute_joint.rs:253-287)
      j.enable_motor = def.enable_motor; j.max_motor_torque = def.max_motor_torque; j.motor_speed = def.motor_speed;
      j.enable_limit = def.enable_limit; j.reference_angle = def.reference_angle;
      j.lower_angle = def.lower_angle; j.upper_angle = def.upper_angle;
    } else if (def.type == J_DISTANCE) {  // b2_distance_joint_new (private b2_distance_joint.rs:43-78)
      j.min_length = b2_max(def.min_length, LINEAR_SLOP);
      j.length = b2_max(def.length, LINEAR_SLOP);
      j.max_length = b2_max(def.max_length, j.min_length);
      j.stiffness = def.stiffness; j.damping = def.damping;
    } else if (def.type == J_PRISMATIC) {  // private joints/b2_prismatic_joint.rs:99-166
      j.local_xaxis_a = def.local_axis_a;
      j.local_xaxis_a.normalize();
      j.local_yaxis_a = b2_cross_sv(1.0f, j.local_xaxis_a);
      j.reference_angle = def.reference_angle;
      j.lower_angle = def.lower_angle; j.upper_angle = def.upper_angle;
      assert(j.lower_angle <= j.upper_angle);
      j.max_motor_torque = def.max_motor_torque; j.motor_speed = def.motor_speed;
      j.enable_limit = def.enable_limit; j.enable_motor = def.enable_motor;
    } else if (def.type == J_FRICTION) {  // B2frictionJoint::new (src/joints/b2_friction_joint.rs:120-150)
      j.max_force = def.max_force; j.max_motor_torque = def.max_motor_torque;
    } else if (def.type == J_PULLEY) {  // B2pulleyJoint::new (private joints/b2_pulley_joint.rs:8-44)
      j.ground_anchor_a = def.ground_anchor_a; j.ground_anchor_b = def.ground_anchor_b;
      j.length = def.length; j.length_b = def.length_b;
      assert(def.ratio != 0.0f);
      j.ratio = def.ratio;
      j.constant = def.length + j.ratio * def.length_b;
    } else if (def.type == J_GEAR) {  // private joints/b2_gear_joint.rs:8-140
      const Joint &j1 = joints[def.joint1], &j2 = joints[def.joint2];
      j.type_a = j1.type; j.type_b = j2.type;
      assert(j.type_a == J_REVOLUTE || j.type_a == J_PRISMATIC);
      assert(j.type_b == J_REVOLUTE || j.type_b == J_PRISMATIC);
      float coordinate_a, coordinate_b;
      // the solver's bodies A / B stay the def's (B2joint::new(&def.base)); the coordinates are measured on the joints' own
      j.body_c = j1.body_a;
      const int jb_a = j1.body_b;
      assert(bodies[jb_a].type == DYNAMIC_BODY);
      const Transform xf_a = bodies[jb_a].xf, xf_c = bodies[j.body_c].xf;
      const float a_a = bodies[jb_a].sweep.a, a_c = bodies[j.body_c].sweep.a;
      j.local_anchor_c = j1.local_anchor_a; j.local_anchor_a = j1.local_anchor_b;
      j.reference_angle = j1.reference_angle;
      if (j.type_a == J_REVOLUTE) {
        j.local_axis_c.set_zero();
        coordinate_a = a_a - a_c - j.reference_angle;
      } else {
        j.local_axis_c = j1.local_xaxis_a;
        Vec2 p_c = j.local_anchor_c;
        Vec2 p_a = b2_mul_t_rot(xf_c.q, b2_mul_rot(xf_a.q, j.local_anchor_a) + (xf_a.p - xf_c.p));
        coordinate_a = b2_dot(p_a - p_c, j.local_axis_c);
      }
      j.body_d = j2.body_a;
      const int jb_b = j2.body_b;
      assert(bodies[jb_b].type == DYNAMIC_BODY);
      const Transform xf_b = bodies[jb_b].xf, xf_d = bodies[j.body_d].xf;
      const float a_b = bodies[jb_b].sweep.a, a_d = bodies[j.body_d].sweep.a;
      j.local_anchor_d = j2.local_anchor_a; j.local_anchor_b = j2.local_anchor_b;
      j.reference_angle_b = j2.reference_angle;
      if (j.type_b == J_REVOLUTE) {
        j.local_axis_d.set_zero();
        coordinate_b = a_b - a_d - j.reference_angle_b;
      } else {
        j.local_axis_d = j2.local_xaxis_a;
        Vec2 p_d = j.local_anchor_d;
        Vec2 p_b = b2_mul_t_rot(xf_d.q, b2_mul_rot(xf_b.q, j.local_anchor_b) + (xf_b.p - xf_d.p));
        coordinate_b = b2_dot(p_b - p_d, j.local_axis_d);
      }
      j.ratio = def.ratio;
      j.constant = coordinate_a + j.ratio * coordinate_b;
    } else if (def.type == J_MOUSE) {  // B2mouseJoint::new (src/joints/b2_mouse_joint.rs:140-170)
      j.ground_anchor_a = def.target;
      j.local_anchor_a.set_zero();  // unused: body A is only the island link
      j.local_anchor_b = b2_mul_t_xf(bodies[def.body_b].xf, def.target);
      j.max_force = def.max_force; j.stiffness = def.stiffness; j.damping = def.damping;
    } else if (def.type == J_MOTOR) {  // B2motorJoint::new (src/joints/b2_motor_joint.rs:175-205): local_anchor_a = linear offset
      j.reference_angle = def.reference_angle;
      j.max_force = def.max_force; j.max_motor_torque = def.max_motor_torque; j.correction_factor = def.correction_factor;
    } else if (def.type == J_WHEEL) {  // B2wheelJoint::new (src/joints/b2_wheel_joint.rs:266-310): the axis is NOT normalised
      j.local_xaxis_a = def.local_axis_a;
      j.local_yaxis_a = b2_cross_sv(1.0f, def.local_axis_a);
      j.lower_angle = def.lower_angle; j.upper_angle = def.upper_angle;
      j.max_motor_torque = def.max_motor_torque; j.motor_speed = def.motor_speed;
      j.enable_limit = def.enable_limit; j.enable_motor = def.enable_motor;
      j.stiffness = def.stiffness; j.damping = def.damping;
    } else if (def.type == J_WELD) {  // B2weldJoint::new (src/joints/b2_weld_joint.rs:152-185)
      j.reference_angle = def.reference_angle;
      j.stiffness = def.stiffness; j.damping = def.damping;
    } else {
      assert(false && "joint type outside the oracle's scope");
    }
    int ji = (int)joints.size();
    joints.push_back(j);
    bodies[def.body_a].joint_edges.push_back(2 * ji);      // edge A on body A's list, then edge B on body B's
    bodies[def.body_b].joint_edges.push_back(2 * ji + 1);
    if (!def.collide_connected)  // flag the contacts between the two bodies for filtering
      for (int e = bodies[def.body_b].contact_list; e != -1; e = edge(e).next)
        if (edge(e).other == def.body_a) contacts[e >> 1].flags |= CF_FILTER;
    return ji;  // creating a joint doesn't wake the bodies
  }
  // B2world::destroy_joint (src/private/dynamics/b2_world.rs:278-339): both bodies wake, the joint leaves the world list and
  // the two body lists (the order of the others is kept), contacts between the bodies are re-filtered when the joint had
  // collide_connected == false.  Joints created after it move down by one index.
  void destroy_joint(int ji) {
    const Joint j = joints[ji];
    set_awake(j.body_a, true);
    set_awake(j.body_b, true);
    joints.erase(joints.begin() + ji);
    for (Body& b : bodies) {
      std::vector<int>& je = b.joint_edges;
      size_t o = 0;
      for (size_t i = 0; i < je.size(); ++i) {
        if ((je[i] >> 1) == ji) continue;
        je[o++] = (je[i] >> 1) > ji ? je[i] - 2 : je[i];
      }
      je.resize(o);
    }
    if (!j.collide_connected)
      for (int e = bodies[j.body_b].contact_list; e != -1; e = edge(e).next)
        if (edge(e).other == j.body_a) contacts[e >> 1].flags |= CF_FILTER;
  }
  // B2mouseJoint::set_target (src/joints/b2_mouse_joint.rs:114-119): wakes body B when the target moves
  void joint_set_target(int ji, Vec2 t) {
    Joint& j = joints[ji];
    if (t.x != j.ground_anchor_a.x || t.y != j.ground_anchor_a.y) { set_awake(j.body_b, true); j.ground_anchor_a = t; }
  }
  // B2revoluteJoint setters (src/joints/b2_revolute_joint.rs:172-242)
  void joint_set_motor_speed(int ji, float speed) {
    Joint& j = joints[ji];
    if (speed != j.motor_speed) { set_awake(j.body_a, true); set_awake(j.body_b, true); j.motor_speed = speed; }
  }
  void joint_set_max_motor_torque(int ji, float torque) {
    Joint& j = joints[ji];
    if (torque != j.max_motor_torque) { set_awake(j.body_a, true); set_awake(j.body_b, true); j.max_motor_torque = torque; }
  }
  void joint_enable_motor(int ji, bool flag) {
    Joint& j = joints[ji];
    if (flag != j.enable_motor) { set_awake(j.body_a, true); set_awake(j.body_b, true); j.enable_motor = flag; }
  }
  void joint_enable_limit(int ji, bool flag) {
    Joint& j = joints[ji];
    if (flag != j.enable_limit) {
      set_awake(j.body_a, true); set_awake(j.body_b, true);
      j.enable_limit = flag; j.lower_impulse = 0.0f; j.upper_impulse = 0.0f;
    }
  }
  void joint_set_limits(int ji, float lower, float upper) {
    Joint& j = joints[ji];
    if (lower != j.lower_angle || upper != j.upper_angle) {
      set_awake(j.body_a, true); set_awake(j.body_b, true);
      j.lower_impulse = 0.0f; j.upper_impulse = 0.0f; j.lower_angle = lower; j.upper_angle = upper;
    }
  }
  bool filter_should_collide(int fa, int fb) const {  // b2_world_callbacks.rs(private):6-18
    const Filter& a = fixtures[fa].filter;
    const Filter& b = fixtures[fb].filter;
    if (a.group_index == b.group_index && a.group_index != 0) return a.group_index > 0;
    return (a.mask_bits & b.category_bits) != 0 && (a.category_bits & b.mask_bits) != 0;
  }

  // -------------------------------------------------------------- contacts
  ContactEdge& edge(int e) { return (e & 1) ? contacts[e >> 1].node_b : contacts[e >> 1].node_a; }
  static bool type_pair_registered(int t1, int t2, bool& primary) {  // b2_contact_registers.rs:67-103
    // primary pairs: (circle,circle) (polygon,circle) (polygon,polygon) (edge,circle) (edge,polygon)
    //                (chain,circle) (chain,polygon)
    auto is_primary = [](int a, int b) {
      return (a == E_CIRCLE && b == E_CIRCLE) || (a == E_POLYGON && b == E_CIRCLE) || (a == E_POLYGON && b == E_POLYGON) ||
             (a == E_EDGE && b == E_CIRCLE) || (a == E_EDGE && b == E_POLYGON) || (a == E_CHAIN && b == E_CIRCLE) ||
             (a == E_CHAIN && b == E_POLYGON);
    };
    if (is_primary(t1, t2)) { primary = true; return true; }
    if (is_primary(t2, t1)) { primary = false; return true; }
    return false;
  }
  void add_pair(int proxy_a, int proxy_b) {  // b2_contact_manager.rs(private):178-302
    int fixture_a = proxies[proxy_a].fixture, fixture_b = proxies[proxy_b].fixture;
    int index_a = proxies[proxy_a].child_index, index_b = proxies[proxy_b].child_index;
    int body_a = fixtures[fixture_a].body, body_b = fixtures[fixture_b].body;
    if (body_a == body_b) return;
    for (int e = bodies[body_b].contact_list; e != -1; e = edge(e).next) {
      if (edge(e).other == body_a) {
        const Contact& c = contacts[e >> 1];
        if (c.fixture_a == fixture_a && c.fixture_b == fixture_b && c.index_a == index_a && c.index_b == index_b) return;
        if (c.fixture_a == fixture_b && c.fixture_b == fixture_a && c.index_a == index_b && c.index_b == index_a) return;
      }
    }
    if (!body_should_collide(body_b, body_a)) return;
    if (!filter_should_collide(fixture_a, fixture_b)) return;
    // B2contact::create — b2_contact.rs(private):11-31
    bool primary = true;
    bool ok = type_pair_registered(fixtures[fixture_a].shape.type, fixtures[fixture_b].shape.type, primary);
    assert(ok && "unregistered shape pair: the reference panics here (unwrap)");
    (void)ok;
    if (!primary) { std::swap(fixture_a, fixture_b); std::swap(index_a, index_b); }
    int ci;
    if (!contact_free.empty()) { ci = contact_free.back(); contact_free.pop_back(); }
    else { ci = (int)contacts.size(); contacts.emplace_back(); }
    Contact& c = contacts[ci];
    c = Contact();
    c.alive = true;
    c.flags = CF_ENABLED;  // b2_contact.rs(private):48-99
    c.fixture_a = fixture_a; c.fixture_b = fixture_b; c.index_a = index_a; c.index_b = index_b;
    c.friction = sqrtf(fixtures[fixture_a].friction * fixtures[fixture_b].friction);
    c.restitution = fixtures[fixture_a].restitution > fixtures[fixture_b].restitution ? fixtures[fixture_a].restitution
                                                                                       : fixtures[fixture_b].restitution;
    c.restitution_threshold = fixtures[fixture_a].restitution_threshold < fixtures[fixture_b].restitution_threshold
                                  ? fixtures[fixture_a].restitution_threshold
                                  : fixtures[fixture_b].restitution_threshold;
    body_a = fixtures[fixture_a].body;
    body_b = fixtures[fixture_b].body;
    // world list push_front
    c.prev = -1;
    c.next = contact_list;
    if (contact_list != -1) contacts[contact_list].prev = ci;
    contact_list = ci;
    // body A edge list push_front
    c.node_a.other = body_b;
    c.node_a.prev = -1;
    c.node_a.next = bodies[body_a].contact_list;
    if (bodies[body_a].contact_list != -1) edge(bodies[body_a].contact_list).prev = 2 * ci;
    bodies[body_a].contact_list = 2 * ci;
    // body B
    c.node_b.other = body_a;
    c.node_b.prev = -1;
    c.node_b.next = bodies[body_b].contact_list;
    if (bodies[body_b].contact_list != -1) edge(bodies[body_b].contact_list).prev = 2 * ci + 1;
    bodies[body_b].contact_list = 2 * ci + 1;
    ++contact_count;
    ++stats.created;
  }
  std::vector<ContactEvent> events;  // cleared at the top of step()
  std::vector<PostSolveEvent> post_solve_events;  // likewise
  void destroy_contact(int ci) {  // b2_contact_manager.rs(private):24-78 ; b2_contact.rs(private):33-46
    Contact& c = contacts[ci];
    if (c.flags & CF_TOUCHING) events.push_back({2, c.fixture_a, c.index_a, c.fixture_b, c.index_b});  // :44-49 end_contact
    int body_a = fixtures[c.fixture_a].body, body_b = fixtures[c.fixture_b].body;
    if (c.prev != -1) contacts[c.prev].next = c.next;
    if (c.next != -1) contacts[c.next].prev = c.prev;
    if (contact_list == ci) contact_list = c.next;
    if (c.node_a.prev != -1) edge(c.node_a.prev).next = c.node_a.next;
    if (c.node_a.next != -1) edge(c.node_a.next).prev = c.node_a.prev;
    if (bodies[body_a].contact_list == 2 * ci) bodies[body_a].contact_list = c.node_a.next;
    if (c.node_b.prev != -1) edge(c.node_b.prev).next = c.node_b.next;
    if (c.node_b.next != -1) edge(c.node_b.next).prev = c.node_b.prev;
    if (bodies[body_b].contact_list == 2 * ci + 1) bodies[body_b].contact_list = c.node_b.next;
    if (c.manifold.point_count > 0 && !fixtures[c.fixture_a].is_sensor && !fixtures[c.fixture_b].is_sensor) {
      set_awake(body_a, true);
      set_awake(body_b, true);
    }
    c.alive = false;
    contact_free.push_back(ci);
    --contact_count;
    ++stats.destroyed;
  }
  void evaluate(const Contact& c, Manifold& m, const Transform& xf_a, const Transform& xf_b) {  // contacts/*.rs:45-56
    const Shape& sa = fixtures[c.fixture_a].shape;
    const Shape& sb = fixtures[c.fixture_b].shape;
    Shape edge_tmp;
    const Shape* a = &sa;
    if (sa.type == E_CHAIN) { chain_get_child_edge(sa, edge_tmp, c.index_a); a = &edge_tmp; }
    if (a->type == E_CIRCLE && sb.type == E_CIRCLE) collide_circles(m, *a, xf_a, sb, xf_b);
    else if (a->type == E_POLYGON && sb.type == E_CIRCLE) collide_polygon_and_circle(m, *a, xf_a, sb, xf_b);
    else if (a->type == E_POLYGON && sb.type == E_POLYGON) collide_polygons(m, *a, xf_a, sb, xf_b);
    else if (a->type == E_EDGE && sb.type == E_CIRCLE) collide_edge_and_circle(m, *a, xf_a, sb, xf_b);
    else if (a->type == E_EDGE && sb.type == E_POLYGON) collide_edge_and_polygon(m, *a, xf_a, sb, xf_b);
    else assert(false);
  }
  void contact_update(int ci) {  // b2_contact.rs(private):103-218
    Contact& c = contacts[ci];
    c.flags |= CF_ENABLED;
    Manifold old_manifold = c.manifold;
    bool was_touching = (c.flags & CF_TOUCHING) != 0;
    bool sensor = fixtures[c.fixture_a].is_sensor || fixtures[c.fixture_b].is_sensor;
    int body_a = fixtures[c.fixture_a].body, body_b = fixtures[c.fixture_b].body;
    const Transform xf_a = bodies[body_a].xf, xf_b = bodies[body_b].xf;
    bool touching;
    if (sensor) {  // b2_contact.rs(private):149-163: GJK overlap, sensors don't generate manifolds
      touching = b2_test_overlap_shapes(fixtures[c.fixture_a].shape, c.index_a, fixtures[c.fixture_b].shape, c.index_b, xf_a, xf_b);
      c.manifold.point_count = 0;
    } else {
      Manifold nm;
      evaluate(c, nm, xf_a, xf_b);
      c.manifold = nm;
      touching = c.manifold.point_count > 0;
      for (int i = 0; i < c.manifold.point_count; ++i) {
        ManifoldPoint& mp2 = c.manifold.points[i];
        mp2.normal_impulse = 0.0f;
        mp2.tangent_impulse = 0.0f;
        for (int j = 0; j < old_manifold.point_count; ++j) {
          const ManifoldPoint& mp1 = old_manifold.points[j];
          if (mp1.id == mp2.id) {
            mp2.normal_impulse = mp1.normal_impulse;
            mp2.tangent_impulse = mp1.tangent_impulse;
            break;
          }
        }
      }
      if (touching != was_touching) {
        set_awake(body_a, true);
        set_awake(body_b, true);
      }
    }
    if (touching) c.flags |= CF_TOUCHING; else c.flags &= ~CF_TOUCHING;
    if (!was_touching && touching) events.push_back({1, c.fixture_a, c.index_a, c.fixture_b, c.index_b});  // :205-207
    if (was_touching && !touching) events.push_back({2, c.fixture_a, c.index_a, c.fixture_b, c.index_b});  // :209-211
  }
  void collide() {  // b2_contact_manager.rs(private):83-171 — destroys AFTER the loop (box2d-rs deviation)
    std::vector<int> to_destroy;
    for (int ci = contact_list; ci != -1; ci = contacts[ci].next) {
      Contact& c = contacts[ci];
      int body_a = fixtures[c.fixture_a].body, body_b = fixtures[c.fixture_b].body;
      if (c.flags & CF_FILTER) {
        if (!body_should_collide(body_b, body_a)) { to_destroy.push_back(ci); continue; }
        if (!filter_should_collide(c.fixture_a, c.fixture_b)) { to_destroy.push_back(ci); continue; }
        c.flags &= ~CF_FILTER;
      }
      bool active_a = (bodies[body_a].flags & BF_AWAKE) && bodies[body_a].type != STATIC_BODY;
      bool active_b = (bodies[body_b].flags & BF_AWAKE) && bodies[body_b].type != STATIC_BODY;
      if (!active_a && !active_b) continue;
      int proxy_id_a = proxies[fixtures[c.fixture_a].proxy_first + c.index_a].proxy_id;
      int proxy_id_b = proxies[fixtures[c.fixture_b].proxy_first + c.index_b].proxy_id;
      if (!broad_phase.test_overlap(proxy_id_a, proxy_id_b)) { to_destroy.push_back(ci); continue; }
      contact_update(ci);
    }
    for (int ci : to_destroy) destroy_contact(ci);
  }
  void find_new_contacts() {  // :173-176
    stats.moved += (int)broad_phase.move_buffer.size();
    broad_phase.update_pairs([this](int pa, int pb) { add_pair(pa, pb); });
    stats.pairs += (int)broad_phase.pair_buffer.size();
  }

  // ---------------------------------------------------------------- island
  struct Island {
    std::vector<int> bodies, contacts, joints;
    std::vector<Position> positions;
    std::vector<Velocity> velocities;
    void clear() { bodies.clear(); contacts.clear(); joints.clear(); }
  };
  std::vector<ContactVelocityConstraint> vcs;
  std::vector<ContactPositionConstraint> pcs;

  void solver_new(const Island& is, const TimeStep& step) {  // b2_contact_solver_private.rs:20-110
    size_t count = is.contacts.size();
    vcs.assign(count, ContactVelocityConstraint());
    pcs.assign(count, ContactPositionConstraint());
    for (size_t i = 0; i < count; ++i) {
      const Contact& contact = contacts[is.contacts[i]];
      const Fixture& fa = fixtures[contact.fixture_a];
      const Fixture& fb = fixtures[contact.fixture_b];
      float radius_a = fa.shape.radius, radius_b = fb.shape.radius;
      const Body& body_a = bodies[fa.body];
      const Body& body_b = bodies[fb.body];
      const Manifold& manifold = contact.manifold;
      int point_count = manifold.point_count;
      ContactVelocityConstraint& vc = vcs[i];
      vc.friction = contact.friction;
      vc.restitution = contact.restitution;
      vc.threshold = contact.restitution_threshold;
      vc.tangent_speed = contact.tangent_speed;
      vc.index_a = body_a.island_index;
      vc.index_b = body_b.island_index;
      vc.inv_mass_a = body_a.inv_mass;
      vc.inv_mass_b = body_b.inv_mass;
      vc.inv_ia = body_a.inv_i;
      vc.inv_ib = body_b.inv_i;
      vc.contact_index = (int)i;
      vc.point_count = point_count;
      vc.k.set_zero();
      vc.normal_mass.set_zero();
      ContactPositionConstraint& pc = pcs[i];
      pc.index_a = body_a.island_index;
      pc.index_b = body_b.island_index;
      pc.inv_mass_a = body_a.inv_mass;
      pc.inv_mass_b = body_b.inv_mass;
      pc.local_center_a = body_a.sweep.local_center;
      pc.local_center_b = body_b.sweep.local_center;
      pc.inv_ia = body_a.inv_i;
      pc.inv_ib = body_b.inv_i;
      pc.local_normal = manifold.local_normal;
      pc.local_point = manifold.local_point;
      pc.point_count = point_count;
      pc.radius_a = radius_a;
      pc.radius_b = radius_b;
      pc.type = manifold.type;
      for (int j = 0; j < point_count; ++j) {
        const ManifoldPoint& cp = manifold.points[j];
        VelocityConstraintPoint& vcp = vc.points[j];
        if (step.warm_starting) {
          vcp.normal_impulse = step.dt_ratio * cp.normal_impulse;
          vcp.tangent_impulse = step.dt_ratio * cp.tangent_impulse;
        } else {
          vcp.normal_impulse = 0.0f;
          vcp.tangent_impulse = 0.0f;
        }
        vcp.r_a.set_zero();
        vcp.r_b.set_zero();
        vcp.normal_mass = 0.0f;
        vcp.tangent_mass = 0.0f;
        vcp.velocity_bias = 0.0f;
        pc.local_points[j] = cp.local_point;
      }
    }
  }
  void initialize_velocity_constraints(const Island& is) {  // :113-226
    for (size_t i = 0; i < is.contacts.size(); ++i) {
      ContactVelocityConstraint& vc = vcs[i];
      ContactPositionConstraint& pc = pcs[i];
      float radius_a = pc.radius_a, radius_b = pc.radius_b;
      const Manifold& manifold = contacts[is.contacts[vc.contact_index]].manifold;
      int index_a = vc.index_a, index_b = vc.index_b;
      float m_a = vc.inv_mass_a, m_b = vc.inv_mass_b, i_a = vc.inv_ia, i_b = vc.inv_ib;
      Vec2 local_center_a = pc.local_center_a, local_center_b = pc.local_center_b;
      Vec2 c_a = is.positions[index_a].c;
      float a_a = is.positions[index_a].a;
      Vec2 v_a = is.velocities[index_a].v;
      float w_a = is.velocities[index_a].w;
      Vec2 c_b = is.positions[index_b].c;
      float a_b = is.positions[index_b].a;
      Vec2 v_b = is.velocities[index_b].v;
      float w_b = is.velocities[index_b].w;
      Transform xf_a, xf_b;
      xf_a.q.set(a_a);
      xf_b.q.set(a_b);
      xf_a.p = c_a - b2_mul_rot(xf_a.q, local_center_a);
      xf_b.p = c_b - b2_mul_rot(xf_b.q, local_center_b);
      WorldManifold wm;
      world_manifold_initialize(wm, manifold, xf_a, radius_a, xf_b, radius_b);
      vc.normal = wm.normal;
      int point_count = vc.point_count;
      for (int j = 0; j < point_count; ++j) {
        VelocityConstraintPoint& vcp = vc.points[j];
        vcp.r_a = wm.points[j] - c_a;
        vcp.r_b = wm.points[j] - c_b;
        float rn_a = b2_cross(vcp.r_a, vc.normal);
        float rn_b = b2_cross(vcp.r_b, vc.normal);
        float k_normal = m_a + m_b + i_a * rn_a * rn_a + i_b * rn_b * rn_b;
        vcp.normal_mass = k_normal > 0.0f ? 1.0f / k_normal : 0.0f;
        Vec2 tangent = b2_cross_vs(vc.normal, 1.0f);
        float rt_a = b2_cross(vcp.r_a, tangent);
        float rt_b = b2_cross(vcp.r_b, tangent);
        float k_tangent = m_a + m_b + i_a * rt_a * rt_a + i_b * rt_b * rt_b;
        vcp.tangent_mass = k_tangent > 0.0f ? 1.0f / k_tangent : 0.0f;
        vcp.velocity_bias = 0.0f;
        float v_rel = b2_dot(vc.normal, v_b + b2_cross_sv(w_b, vcp.r_b) - v_a - b2_cross_sv(w_a, vcp.r_a));
        if (v_rel < -vc.threshold) vcp.velocity_bias = -vc.restitution * v_rel;
      }
      if (vc.point_count == 2 && block_solve) {
        const VelocityConstraintPoint& vcp1 = vc.points[0];
        const VelocityConstraintPoint& vcp2 = vc.points[1];
        float rn1_a = b2_cross(vcp1.r_a, vc.normal);
        float rn1_b = b2_cross(vcp1.r_b, vc.normal);
        float rn2_a = b2_cross(vcp2.r_a, vc.normal);
        float rn2_b = b2_cross(vcp2.r_b, vc.normal);
        float k11 = m_a + m_b + i_a * rn1_a * rn1_a + i_b * rn1_b * rn1_b;
        float k22 = m_a + m_b + i_a * rn2_a * rn2_a + i_b * rn2_b * rn2_b;
        float k12 = m_a + m_b + i_a * rn1_a * rn2_a + i_b * rn1_b * rn2_b;
        const float k_max_condition_number = 1000.0f;
        if (k11 * k11 < k_max_condition_number * (k11 * k22 - k12 * k12)) {
          vc.k.ex.set(k11, k12);
          vc.k.ey.set(k12, k22);
          vc.normal_mass = vc.k.get_inverse();
        } else {
          vc.point_count = 1;
        }
      }
    }
  }
  void warm_start(Island& is) {  // :228-266
    for (auto& vc : vcs) {
      int index_a = vc.index_a, index_b = vc.index_b;
      float m_a = vc.inv_mass_a, i_a = vc.inv_ia, m_b = vc.inv_mass_b, i_b = vc.inv_ib;
      int point_count = vc.point_count;
      Vec2 v_a = is.velocities[index_a].v;
      float w_a = is.velocities[index_a].w;
      Vec2 v_b = is.velocities[index_b].v;
      float w_b = is.velocities[index_b].w;
      Vec2 normal = vc.normal;
      Vec2 tangent = b2_cross_vs(normal, 1.0f);
      for (int j = 0; j < point_count; ++j) {
        const VelocityConstraintPoint& vcp = vc.points[j];
        Vec2 p = vcp.normal_impulse * normal + vcp.tangent_impulse * tangent;
        w_a -= i_a * b2_cross(vcp.r_a, p);
        v_a -= m_a * p;
        w_b += i_b * b2_cross(vcp.r_b, p);
        v_b += m_b * p;
      }
      is.velocities[index_a].v = v_a;
      is.velocities[index_a].w = w_a;
      is.velocities[index_b].v = v_b;
      is.velocities[index_b].w = w_b;
    }
  }
  void solve_velocity_constraints(Island& is) {  // :268-583
    for (auto& vc : vcs) {
      int index_a = vc.index_a, index_b = vc.index_b;
      float m_a = vc.inv_mass_a, i_a = vc.inv_ia, m_b = vc.inv_mass_b, i_b = vc.inv_ib;
      int point_count = vc.point_count;
      Vec2 v_a = is.velocities[index_a].v;
      float w_a = is.velocities[index_a].w;
      Vec2 v_b = is.velocities[index_b].v;
      float w_b = is.velocities[index_b].w;
      Vec2 normal = vc.normal;
      Vec2 tangent = b2_cross_vs(normal, 1.0f);
      float friction = vc.friction;
      for (int j = 0; j < point_count; ++j) {
        VelocityConstraintPoint& vcp = vc.points[j];
        Vec2 dv = v_b + b2_cross_sv(w_b, vcp.r_b) - v_a - b2_cross_sv(w_a, vcp.r_a);
        float vt = b2_dot(dv, tangent) - vc.tangent_speed;
        float lambda = vcp.tangent_mass * (-vt);
        float max_friction = friction * vcp.normal_impulse;
        float new_impulse = b2_clamp(vcp.tangent_impulse + lambda, -max_friction, max_friction);
        lambda = new_impulse - vcp.tangent_impulse;
        vcp.tangent_impulse = new_impulse;
        Vec2 p = lambda * tangent;
        v_a -= m_a * p;
        w_a -= i_a * b2_cross(vcp.r_a, p);
        v_b += m_b * p;
        w_b += i_b * b2_cross(vcp.r_b, p);
      }
      if (point_count == 1 || block_solve == false) {
        for (int j = 0; j < point_count; ++j) {
          VelocityConstraintPoint& vcp = vc.points[j];
          Vec2 dv = v_b + b2_cross_sv(w_b, vcp.r_b) - v_a - b2_cross_sv(w_a, vcp.r_a);
          float vn = b2_dot(dv, normal);
          float lambda = -vcp.normal_mass * (vn - vcp.velocity_bias);
          float new_impulse = b2_max(vcp.normal_impulse + lambda, 0.0f);
          lambda = new_impulse - vcp.normal_impulse;
          vcp.normal_impulse = new_impulse;
          Vec2 p = lambda * normal;
          v_a -= m_a * p;
          w_a -= i_a * b2_cross(vcp.r_a, p);
          v_b += m_b * p;
          w_b += i_b * b2_cross(vcp.r_b, p);
        }
      } else {
        VelocityConstraintPoint& cp1 = vc.points[0];
        VelocityConstraintPoint& cp2 = vc.points[1];
        Vec2 a(cp1.normal_impulse, cp2.normal_impulse);
        Vec2 dv1 = v_b + b2_cross_sv(w_b, cp1.r_b) - v_a - b2_cross_sv(w_a, cp1.r_a);
        Vec2 dv2 = v_b + b2_cross_sv(w_b, cp2.r_b) - v_a - b2_cross_sv(w_a, cp2.r_a);
        float vn1 = b2_dot(dv1, normal);
        float vn2 = b2_dot(dv2, normal);
        Vec2 b(vn1 - cp1.velocity_bias, vn2 - cp2.velocity_bias);
        b -= b2_mul(vc.k, a);
        for (;;) {
          Vec2 x = -b2_mul(vc.normal_mass, b);
          if (x.x >= 0.0f && x.y >= 0.0f) {
            Vec2 d = x - a;
            Vec2 p1 = d.x * normal, p2 = d.y * normal;
            v_a -= m_a * (p1 + p2);
            w_a -= i_a * (b2_cross(cp1.r_a, p1) + b2_cross(cp2.r_a, p2));
            v_b += m_b * (p1 + p2);
            w_b += i_b * (b2_cross(cp1.r_b, p1) + b2_cross(cp2.r_b, p2));
            cp1.normal_impulse = x.x;
            cp2.normal_impulse = x.y;
            break;
          }
          x.x = -cp1.normal_mass * b.x;
          x.y = 0.0f;
          vn1 = 0.0f;
          vn2 = vc.k.ex.y * x.x + b.y;
          if (x.x >= 0.0f && vn2 >= 0.0f) {
            Vec2 d = x - a;
            Vec2 p1 = d.x * normal, p2 = d.y * normal;
            v_a -= m_a * (p1 + p2);
            w_a -= i_a * (b2_cross(cp1.r_a, p1) + b2_cross(cp2.r_a, p2));
            v_b += m_b * (p1 + p2);
            w_b += i_b * (b2_cross(cp1.r_b, p1) + b2_cross(cp2.r_b, p2));
            cp1.normal_impulse = x.x;
            cp2.normal_impulse = x.y;
            break;
          }
          x.x = 0.0f;
          x.y = -cp2.normal_mass * b.y;
          vn1 = vc.k.ey.x * x.y + b.x;
          vn2 = 0.0f;
          if (x.y >= 0.0f && vn1 >= 0.0f) {
            Vec2 d = x - a;
            Vec2 p1 = d.x * normal, p2 = d.y * normal;
            v_a -= m_a * (p1 + p2);
            w_a -= i_a * (b2_cross(cp1.r_a, p1) + b2_cross(cp2.r_a, p2));
            v_b += m_b * (p1 + p2);
            w_b += i_b * (b2_cross(cp1.r_b, p1) + b2_cross(cp2.r_b, p2));
            cp1.normal_impulse = x.x;
            cp2.normal_impulse = x.y;
            break;
          }
          x.x = 0.0f;
          x.y = 0.0f;
          vn1 = b.x;
          vn2 = b.y;
          if (vn1 >= 0.0f && vn2 >= 0.0f) {
            Vec2 d = x - a;
            Vec2 p1 = d.x * normal, p2 = d.y * normal;
            v_a -= m_a * (p1 + p2);
            w_a -= i_a * (b2_cross(cp1.r_a, p1) + b2_cross(cp2.r_a, p2));
            v_b += m_b * (p1 + p2);
            w_b += i_b * (b2_cross(cp1.r_b, p1) + b2_cross(cp2.r_b, p2));
            cp1.normal_impulse = x.x;
            cp2.normal_impulse = x.y;
            break;
          }
          break;
        }
      }
      is.velocities[index_a].v = v_a;
      is.velocities[index_a].w = w_a;
      is.velocities[index_b].v = v_b;
      is.velocities[index_b].w = w_b;
    }
  }
  void store_impulses(const Island& is) {  // :585-598
    for (auto& vc : vcs) {
      Manifold& manifold = contacts[is.contacts[vc.contact_index]].manifold;
      for (int j = 0; j < vc.point_count; ++j) {
        manifold.points[j].normal_impulse = vc.points[j].normal_impulse;
        manifold.points[j].tangent_impulse = vc.points[j].tangent_impulse;
      }
    }
  }
  bool solve_position_constraints(Island& is) {  // :600-730
    float min_separation = 0.0f;
    for (auto& pc : pcs) {
      int index_a = pc.index_a, index_b = pc.index_b;
      Vec2 local_center_a = pc.local_center_a;
      float m_a = pc.inv_mass_a, i_a = pc.inv_ia;
      Vec2 local_center_b = pc.local_center_b;
      float m_b = pc.inv_mass_b, i_b = pc.inv_ib;
      int point_count = pc.point_count;
      Vec2 c_a = is.positions[index_a].c;
      float a_a = is.positions[index_a].a;
      Vec2 c_b = is.positions[index_b].c;
      float a_b = is.positions[index_b].a;
      for (int j = 0; j < point_count; ++j) {
        Transform xf_a, xf_b;
        xf_a.q.set(a_a);
        xf_b.q.set(a_b);
        xf_a.p = c_a - b2_mul_rot(xf_a.q, local_center_a);
        xf_b.p = c_b - b2_mul_rot(xf_b.q, local_center_b);
        Vec2 normal, point;
        float separation;
        switch (pc.type) {
          case E_CIRCLES: {
            Vec2 point_a = b2_mul_xf(xf_a, pc.local_point);
            Vec2 point_b = b2_mul_xf(xf_b, pc.local_points[0]);
            normal = point_b - point_a;
            normal.normalize();
            point = 0.5f * (point_a + point_b);
            separation = b2_dot(point_b - point_a, normal) - pc.radius_a - pc.radius_b;
          } break;
          case E_FACE_A: {
            normal = b2_mul_rot(xf_a.q, pc.local_normal);
            Vec2 plane_point = b2_mul_xf(xf_a, pc.local_point);
            Vec2 clip_point = b2_mul_xf(xf_b, pc.local_points[j]);
            separation = b2_dot(clip_point - plane_point, normal) - pc.radius_a - pc.radius_b;
            point = clip_point;
          } break;
          default: {
            normal = b2_mul_rot(xf_b.q, pc.local_normal);
            Vec2 plane_point = b2_mul_xf(xf_b, pc.local_point);
            Vec2 clip_point = b2_mul_xf(xf_a, pc.local_points[j]);
            separation = b2_dot(clip_point - plane_point, normal) - pc.radius_a - pc.radius_b;
            point = clip_point;
            normal = -normal;
          } break;
        }
        Vec2 r_a = point - c_a, r_b = point - c_b;
        min_separation = b2_min(min_separation, separation);
        float c = b2_clamp(BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
        float rn_a = b2_cross(r_a, normal);
        float rn_b = b2_cross(r_b, normal);
        float k = m_a + m_b + i_a * rn_a * rn_a + i_b * rn_b * rn_b;
        float impulse = k > 0.0f ? -c / k : 0.0f;
        Vec2 p = impulse * normal;
        c_a -= m_a * p;
        a_a -= i_a * b2_cross(r_a, p);
        c_b += m_b * p;
        a_b += i_b * b2_cross(r_b, p);
      }
      is.positions[index_a].c = c_a;
      is.positions[index_a].a = a_a;
      is.positions[index_b].c = c_b;
      is.positions[index_b].a = a_b;
    }
    return min_separation >= -3.0f * LINEAR_SLOP;
  }

  // -------------------------------------------------------------- joint solver
  // B2jointTraitDyn::init_velocity_constraints / solve_velocity_constraints / solve_position_constraints
  // (src/b2_joint.rs:268-286), dispatched on the joint type.
  // The gear joint couples four bodies (private joints/b2_gear_joint.rs:142-245 / :247-290 / :292-400).  Every body is read
  // first and written back in the order A, B, C, D, as the reference does: when two of them are one body (both joints on one
  // carrier, or joint 2 hanging off body A) the later store wins.
  struct GearRows { Vec2 jv_ac, jv_bd; float jw_a, jw_b, jw_c, jw_d, mass, coordinate_a, coordinate_b; };
  GearRows gear_rows(const Joint& j, const Vec2 c[4], const float a[4], const Vec2 lc[4], const float m[4], const float i[4]) const {
    GearRows g;
    Rot q_a(a[0]), q_b(a[1]), q_c(a[2]), q_d(a[3]);
    g.mass = 0.0f;
    if (j.type_a == J_REVOLUTE) {
      g.jv_ac.set_zero();
      g.jw_a = 1.0f; g.jw_c = 1.0f;
      g.mass += i[0] + i[2];
      g.coordinate_a = a[0] - a[2] - j.reference_angle;
    } else {
      Vec2 u = b2_mul_rot(q_c, j.local_axis_c);
      Vec2 r_c = b2_mul_rot(q_c, j.local_anchor_c - lc[2]);
      Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - lc[0]);
      g.jv_ac = u;
      g.jw_c = b2_cross(r_c, u);
      g.jw_a = b2_cross(r_a, u);
      g.mass += m[2] + m[0] + i[2] * g.jw_c * g.jw_c + i[0] * g.jw_a * g.jw_a;
      Vec2 p_c = j.local_anchor_c - lc[2];
      Vec2 p_a = b2_mul_t_rot(q_c, r_a + (c[0] - c[2]));
      g.coordinate_a = b2_dot(p_a - p_c, j.local_axis_c);
    }
    if (j.type_b == J_REVOLUTE) {
      g.jv_bd.set_zero();
      g.jw_b = j.ratio; g.jw_d = j.ratio;
      g.mass += j.ratio * j.ratio * (i[1] + i[3]);
      g.coordinate_b = a[1] - a[3] - j.reference_angle_b;
    } else {
      Vec2 u = b2_mul_rot(q_d, j.local_axis_d);
      Vec2 r_d = b2_mul_rot(q_d, j.local_anchor_d - lc[3]);
      Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - lc[1]);
      g.jv_bd = j.ratio * u;
      g.jw_d = j.ratio * b2_cross(r_d, u);
      g.jw_b = j.ratio * b2_cross(r_b, u);
      g.mass += j.ratio * j.ratio * (m[3] + m[1]) + i[3] * g.jw_d * g.jw_d + i[1] * g.jw_b * g.jw_b;
      Vec2 p_d = j.local_anchor_d - lc[3];
      Vec2 p_b = b2_mul_t_rot(q_d, r_b + (c[1] - c[3]));
      g.coordinate_b = b2_dot(p_b - p_d, j.local_axis_d);
    }
    return g;
  }
  void gear_bodies(const Joint& j, int ix[4], Vec2 lc[4], float m[4], float i[4]) const {
    const int b[4] = {j.body_a, j.body_b, j.body_c, j.body_d};
    for (int k = 0; k < 4; ++k) {
      ix[k] = bodies[b[k]].island_index; lc[k] = bodies[b[k]].sweep.local_center;
      m[k] = bodies[b[k]].inv_mass; i[k] = bodies[b[k]].inv_i;
    }
  }
  void gear_init_velocity(Joint& j, const TimeStep& step, Island& is) {
    int ix[4]; Vec2 lc[4], c[4], v[4]; float m[4], i[4], a[4], w[4];
    gear_bodies(j, ix, lc, m, i);
    for (int k = 0; k < 4; ++k) { c[k] = is.positions[ix[k]].c; a[k] = is.positions[ix[k]].a; v[k] = is.velocities[ix[k]].v; w[k] = is.velocities[ix[k]].w; }
    const GearRows g = gear_rows(j, c, a, lc, m, i);
    j.jv_ac = g.jv_ac; j.jv_bd = g.jv_bd; j.jw_a = g.jw_a; j.jw_b = g.jw_b; j.jw_c = g.jw_c; j.jw_d = g.jw_d;
    j.mass = g.mass > 0.0f ? 1.0f / g.mass : 0.0f;
    if (step.warm_starting) {  // not rescaled by dt_ratio
      v[0] += (m[0] * j.impulse) * j.jv_ac;
      w[0] += i[0] * j.impulse * j.jw_a;
      v[1] += (m[1] * j.impulse) * j.jv_bd;
      w[1] += i[1] * j.impulse * j.jw_b;
      v[2] -= (m[2] * j.impulse) * j.jv_ac;
      w[2] -= i[2] * j.impulse * j.jw_c;
      v[3] -= (m[3] * j.impulse) * j.jv_bd;
      w[3] -= i[3] * j.impulse * j.jw_d;
    } else {
      j.impulse = 0.0f;
    }
    for (int k = 0; k < 4; ++k) { is.velocities[ix[k]].v = v[k]; is.velocities[ix[k]].w = w[k]; }
  }
  void gear_solve_velocity(Joint& j, Island& is) {
    int ix[4]; Vec2 lc[4], v[4]; float m[4], i[4], w[4];
    gear_bodies(j, ix, lc, m, i);
    for (int k = 0; k < 4; ++k) { v[k] = is.velocities[ix[k]].v; w[k] = is.velocities[ix[k]].w; }
    float cdot = b2_dot(j.jv_ac, v[0] - v[2]) + b2_dot(j.jv_bd, v[1] - v[3]);
    cdot += (j.jw_a * w[0] - j.jw_c * w[2]) + (j.jw_b * w[1] - j.jw_d * w[3]);
    float impulse = -j.mass * cdot;
    j.impulse += impulse;
    v[0] += (m[0] * impulse) * j.jv_ac;
    w[0] += i[0] * impulse * j.jw_a;
    v[1] += (m[1] * impulse) * j.jv_bd;
    w[1] += i[1] * impulse * j.jw_b;
    v[2] -= (m[2] * impulse) * j.jv_ac;
    w[2] -= i[2] * impulse * j.jw_c;
    v[3] -= (m[3] * impulse) * j.jv_bd;
    w[3] -= i[3] * impulse * j.jw_d;
    for (int k = 0; k < 4; ++k) { is.velocities[ix[k]].v = v[k]; is.velocities[ix[k]].w = w[k]; }
  }
  bool gear_solve_position(Joint& j, Island& is) {
    int ix[4]; Vec2 lc[4], c[4]; float m[4], i[4], a[4];
    gear_bodies(j, ix, lc, m, i);
    for (int k = 0; k < 4; ++k) { c[k] = is.positions[ix[k]].c; a[k] = is.positions[ix[k]].a; }
    const GearRows g = gear_rows(j, c, a, lc, m, i);
    float cc = (g.coordinate_a + j.ratio * g.coordinate_b) - j.constant;
    float impulse = 0.0f;
    if (g.mass > 0.0f) impulse = -cc / g.mass;
    c[0] += m[0] * impulse * g.jv_ac;
    a[0] += i[0] * impulse * g.jw_a;
    c[1] += m[1] * impulse * g.jv_bd;
    a[1] += i[1] * impulse * g.jw_b;
    c[2] -= m[2] * impulse * g.jv_ac;
    a[2] -= i[2] * impulse * g.jw_c;
    c[3] -= m[3] * impulse * g.jv_bd;
    a[3] -= i[3] * impulse * g.jw_d;
    for (int k = 0; k < 4; ++k) { is.positions[ix[k]].c = c[k]; is.positions[ix[k]].a = a[k]; }
    return true;  // linear_error stays 0
  }
  void joint_init_velocity_constraints(Joint& j, const TimeStep& step, Island& is) {
    if (j.type == J_GEAR) { gear_init_velocity(j, step, is); return; }
    const Body& body_a = bodies[j.body_a];
    const Body& body_b = bodies[j.body_b];
    j.index_a = body_a.island_index;
    j.index_b = body_b.island_index;
    j.local_center_a = body_a.sweep.local_center;
    j.local_center_b = body_b.sweep.local_center;
    j.inv_mass_a = body_a.inv_mass;
    j.inv_mass_b = body_b.inv_mass;
    j.inv_ia = body_a.inv_i;
    j.inv_ib = body_b.inv_i;
    Vec2 c_a = is.positions[j.index_a].c;
    float a_a = is.positions[j.index_a].a;
    Vec2 v_a = is.velocities[j.index_a].v;
    float w_a = is.velocities[j.index_a].w;
    Vec2 c_b = is.positions[j.index_b].c;
    float a_b = is.positions[j.index_b].a;
    Vec2 v_b = is.velocities[j.index_b].v;
    float w_b = is.velocities[j.index_b].w;
    Rot q_a(a_a), q_b(a_b);
    j.r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
    j.r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
    if (j.type == J_REVOLUTE) {  // private joints/b2_revolute_joint.rs:22-123
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      j.k.ex.x = m_a + m_b + j.r_a.y * j.r_a.y * i_a + j.r_b.y * j.r_b.y * i_b;
      j.k.ey.x = -j.r_a.y * j.r_a.x * i_a - j.r_b.y * j.r_b.x * i_b;
      j.k.ex.y = j.k.ey.x;
      j.k.ey.y = m_a + m_b + j.r_a.x * j.r_a.x * i_a + j.r_b.x * j.r_b.x * i_b;
      j.axial_mass = i_a + i_b;
      bool fixed_rotation;
      if (j.axial_mass > 0.0f) { j.axial_mass = 1.0f / j.axial_mass; fixed_rotation = false; }
      else fixed_rotation = true;
      j.angle = a_b - a_a - j.reference_angle;
      if (j.enable_limit == false || fixed_rotation) { j.lower_impulse = 0.0f; j.upper_impulse = 0.0f; }
      if (j.enable_motor == false || fixed_rotation) j.motor_impulse = 0.0f;
      if (step.warm_starting) {
        j.impulse2 *= step.dt_ratio;
        j.motor_impulse *= step.dt_ratio;
        j.lower_impulse *= step.dt_ratio;
        j.upper_impulse *= step.dt_ratio;
        float axial_impulse = j.motor_impulse + j.lower_impulse - j.upper_impulse;
        Vec2 p(j.impulse2.x, j.impulse2.y);
        v_a -= m_a * p;
        w_a -= i_a * (b2_cross(j.r_a, p) + axial_impulse);
        v_b += m_b * p;
        w_b += i_b * (b2_cross(j.r_b, p) + axial_impulse);
      } else {
        j.impulse2.set_zero();
        j.motor_impulse = 0.0f;
        j.lower_impulse = 0.0f;
        j.upper_impulse = 0.0f;
      }
    } else if (j.type == J_PRISMATIC) {  // private joints/b2_prismatic_joint.rs:168-280
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      Vec2 d = (c_b - c_a) + j.r_b - j.r_a;
      {
        j.axis = b2_mul_rot(q_a, j.local_xaxis_a);
        j.a1 = b2_cross(d + j.r_a, j.axis);
        j.a2 = b2_cross(j.r_b, j.axis);
        j.axial_mass = m_a + m_b + i_a * j.a1 * j.a1 + i_b * j.a2 * j.a2;
        if (j.axial_mass > 0.0f) j.axial_mass = 1.0f / j.axial_mass;
      }
      {
        j.perp = b2_mul_rot(q_a, j.local_yaxis_a);
        j.s1 = b2_cross(d + j.r_a, j.perp);
        j.s2 = b2_cross(j.r_b, j.perp);
        float k11 = m_a + m_b + i_a * j.s1 * j.s1 + i_b * j.s2 * j.s2;
        float k12 = i_a * j.s1 + i_b * j.s2;
        float k22 = i_a + i_b;
        if (k22 == 0.0f) k22 = 1.0f;  // bodies with fixed rotation
        j.k.ex.set(k11, k12);
        j.k.ey.set(k12, k22);
      }
      if (j.enable_limit) {
        j.translation = b2_dot(j.axis, d);
      } else {
        j.lower_impulse = 0.0f;
        j.upper_impulse = 0.0f;
      }
      if (j.enable_motor == false) j.motor_impulse = 0.0f;
      if (step.warm_starting) {
        j.impulse2 *= step.dt_ratio;
        j.motor_impulse *= step.dt_ratio;
        j.lower_impulse *= step.dt_ratio;
        j.upper_impulse *= step.dt_ratio;
        float axial_impulse = j.motor_impulse + j.lower_impulse - j.upper_impulse;
        Vec2 p = j.impulse2.x * j.perp + axial_impulse * j.axis;
        float la = j.impulse2.x * j.s1 + j.impulse2.y + axial_impulse * j.a1;
        float lb = j.impulse2.x * j.s2 + j.impulse2.y + axial_impulse * j.a2;
        v_a -= m_a * p;
        w_a -= i_a * la;
        v_b += m_b * p;
        w_b += i_b * lb;
      } else {
        j.impulse2.set_zero();
        j.motor_impulse = 0.0f;
        j.lower_impulse = 0.0f;
        j.upper_impulse = 0.0f;
      }
    } else if (j.type == J_PULLEY) {  // private joints/b2_pulley_joint.rs:46-137
      j.u = c_a + j.r_a - j.ground_anchor_a;
      j.u_b = c_b + j.r_b - j.ground_anchor_b;
      float length_a = j.u.length(), length_b = j.u_b.length();
      if (length_a > 10.0f * LINEAR_SLOP) j.u *= 1.0f / length_a; else j.u.set_zero();
      if (length_b > 10.0f * LINEAR_SLOP) j.u_b *= 1.0f / length_b; else j.u_b.set_zero();
      float ru_a = b2_cross(j.r_a, j.u), ru_b = b2_cross(j.r_b, j.u_b);
      float m_a = j.inv_mass_a + j.inv_ia * ru_a * ru_a;
      float m_b = j.inv_mass_b + j.inv_ib * ru_b * ru_b;
      j.mass = m_a + j.ratio * j.ratio * m_b;
      if (j.mass > 0.0f) j.mass = 1.0f / j.mass;
      if (step.warm_starting) {
        j.impulse *= step.dt_ratio;
        Vec2 pa = -(j.impulse) * j.u;
        Vec2 pb = (-j.ratio * j.impulse) * j.u_b;
        v_a += j.inv_mass_a * pa;
        w_a += j.inv_ia * b2_cross(j.r_a, pa);
        v_b += j.inv_mass_b * pb;
        w_b += j.inv_ib * b2_cross(j.r_b, pb);
      } else {
        j.impulse = 0.0f;
      }
    } else if (j.type == J_MOUSE) {  // private joints/b2_mouse_joint.rs:7-67: body A is not read or written
      float d = j.damping, k = j.stiffness, h = step.dt;
      j.gamma = h * (d + h * k);
      if (j.gamma != 0.0f) j.gamma = 1.0f / j.gamma;
      j.beta = h * k * j.gamma;
      Mat22 km;
      km.ex.x = j.inv_mass_b + j.inv_ib * j.r_b.y * j.r_b.y + j.gamma;
      km.ex.y = -j.inv_ib * j.r_b.x * j.r_b.y;
      km.ey.x = km.ex.y;
      km.ey.y = j.inv_mass_b + j.inv_ib * j.r_b.x * j.r_b.x + j.gamma;
      j.k = km.get_inverse();
      j.linear_error = c_b + j.r_b - j.ground_anchor_a;
      j.linear_error *= j.beta;
      w_b *= 0.98f;
      if (step.warm_starting) {
        j.impulse2 *= step.dt_ratio;
        v_b += j.inv_mass_b * j.impulse2;
        w_b += j.inv_ib * b2_cross(j.r_b, j.impulse2);
      } else {
        j.impulse2.set_zero();
      }
    } else if (j.type == J_FRICTION || j.type == J_MOTOR) {
      // private joints/b2_friction_joint.rs:8-72, b2_motor_joint.rs:8-96: the same rows; the motor joint measures from body
      // B's origin to the linear offset on body A and carries position errors
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      if (j.type == J_MOTOR) {
        j.r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
        j.r_b = b2_mul_rot(q_b, -j.local_center_b);
      }
      Mat22 k;
      k.ex.x = m_a + m_b + i_a * j.r_a.y * j.r_a.y + i_b * j.r_b.y * j.r_b.y;
      k.ex.y = -i_a * j.r_a.x * j.r_a.y - i_b * j.r_b.x * j.r_b.y;
      k.ey.x = k.ex.y;
      k.ey.y = m_a + m_b + i_a * j.r_a.x * j.r_a.x + i_b * j.r_b.x * j.r_b.x;
      j.k = k.get_inverse();
      j.axial_mass = i_a + i_b;
      if (j.axial_mass > 0.0f) j.axial_mass = 1.0f / j.axial_mass;
      if (j.type == J_MOTOR) {
        j.linear_error = c_b + j.r_b - c_a - j.r_a;
        j.angular_error = a_b - a_a - j.reference_angle;
      }
      if (step.warm_starting) {
        j.impulse2 *= step.dt_ratio;
        j.motor_impulse *= step.dt_ratio;
        Vec2 p(j.impulse2.x, j.impulse2.y);
        v_a -= m_a * p;
        w_a -= i_a * (b2_cross(j.r_a, p) + j.motor_impulse);
        v_b += m_b * p;
        w_b += i_b * (b2_cross(j.r_b, p) + j.motor_impulse);
      } else {
        j.impulse2.set_zero();
        j.motor_impulse = 0.0f;
      }
    } else if (j.type == J_WHEEL) {  // private joints/b2_wheel_joint.rs:19-170
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      Vec2 d = c_b + j.r_b - c_a - j.r_a;
      {  // point to line constraint
        j.ay = b2_mul_rot(q_a, j.local_yaxis_a);
        j.s_ay = b2_cross(d + j.r_a, j.ay);
        j.s_by = b2_cross(j.r_b, j.ay);
        j.mass = m_a + m_b + i_a * j.s_ay * j.s_ay + i_b * j.s_by * j.s_by;
        if (j.mass > 0.0f) j.mass = 1.0f / j.mass;
      }
      // spring constraint
      j.ax = b2_mul_rot(q_a, j.local_xaxis_a);
      j.s_ax = b2_cross(d + j.r_a, j.ax);
      j.s_bx = b2_cross(j.r_b, j.ax);
      float inv_mass = m_a + m_b + i_a * j.s_ax * j.s_ax + i_b * j.s_bx * j.s_bx;
      if (inv_mass > 0.0f) j.axial_mass = 1.0f / inv_mass;
      else j.axial_mass = 0.0f;
      j.spring_mass = 0.0f;
      j.bias = 0.0f;
      j.gamma = 0.0f;
      if (j.stiffness > 0.0f && inv_mass > 0.0f) {
        j.spring_mass = 1.0f / inv_mass;
        float c = b2_dot(d, j.ax);
        float h = step.dt;
        j.gamma = h * (j.damping + h * j.stiffness);
        if (j.gamma > 0.0f) j.gamma = 1.0f / j.gamma;
        j.bias = c * h * j.stiffness * j.gamma;
        j.spring_mass = inv_mass + j.gamma;
        if (j.spring_mass > 0.0f) j.spring_mass = 1.0f / j.spring_mass;
      } else {
        j.spring_impulse = 0.0f;
      }
      if (j.enable_limit) {
        j.translation = b2_dot(j.ax, d);
      } else {
        j.lower_impulse = 0.0f;
        j.upper_impulse = 0.0f;
      }
      if (j.enable_motor) {
        j.motor_mass = i_a + i_b;
        if (j.motor_mass > 0.0f) j.motor_mass = 1.0f / j.motor_mass;
      } else {
        j.motor_mass = 0.0f;
        j.motor_impulse = 0.0f;
      }
      if (step.warm_starting) {
        // account for variable time step (the limit impulses are not scaled: b2_wheel_joint.rs:143-146)
        j.impulse *= step.dt_ratio;
        j.spring_impulse *= step.dt_ratio;
        j.motor_impulse *= step.dt_ratio;
        float axial_impulse = j.spring_impulse + j.lower_impulse - j.upper_impulse;
        Vec2 p = j.impulse * j.ay + axial_impulse * j.ax;
        float la = j.impulse * j.s_ay + axial_impulse * j.s_ax + j.motor_impulse;
        float lb = j.impulse * j.s_by + axial_impulse * j.s_bx + j.motor_impulse;
        v_a -= j.inv_mass_a * p;
        w_a -= j.inv_ia * la;
        v_b += j.inv_mass_b * p;
        w_b += j.inv_ib * lb;
      } else {
        j.impulse = 0.0f;
        j.spring_impulse = 0.0f;
        j.motor_impulse = 0.0f;
        j.lower_impulse = 0.0f;
        j.upper_impulse = 0.0f;
      }
    } else if (j.type == J_WELD) {  // private joints/b2_weld_joint.rs:22-136
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      float k[9];  // ex.x ex.y ex.z ey.x ey.y ey.z ez.x ez.y ez.z
      k[0] = m_a + m_b + j.r_a.y * j.r_a.y * i_a + j.r_b.y * j.r_b.y * i_b;
      k[3] = -j.r_a.y * j.r_a.x * i_a - j.r_b.y * j.r_b.x * i_b;
      k[6] = -j.r_a.y * i_a - j.r_b.y * i_b;
      k[1] = k[3];
      k[4] = m_a + m_b + j.r_a.x * j.r_a.x * i_a + j.r_b.x * j.r_b.x * i_b;
      k[7] = j.r_a.x * i_a + j.r_b.x * i_b;
      k[2] = k[6];
      k[5] = k[7];
      k[8] = i_a + i_b;
      if (j.stiffness > 0.0f) {
        mat33_get_inverse22(k, j.m33);
        float inv_m = i_a + i_b;
        float c = a_b - a_a - j.reference_angle;
        float d = j.damping, kk = j.stiffness, h = step.dt;
        j.gamma = h * (d + h * kk);
        j.gamma = j.gamma != 0.0f ? 1.0f / j.gamma : 0.0f;
        j.bias = c * h * kk * j.gamma;
        inv_m += j.gamma;
        j.m33[8] = inv_m != 0.0f ? 1.0f / inv_m : 0.0f;
      } else if (k[8] == 0.0f) {
        mat33_get_inverse22(k, j.m33);
        j.gamma = 0.0f;
        j.bias = 0.0f;
      } else {
        mat33_get_sym_inverse33(k, j.m33);
        j.gamma = 0.0f;
        j.bias = 0.0f;
      }
      if (step.warm_starting) {
        j.impulse3[0] *= step.dt_ratio; j.impulse3[1] *= step.dt_ratio; j.impulse3[2] *= step.dt_ratio;
        Vec2 p(j.impulse3[0], j.impulse3[1]);
        v_a -= m_a * p;
        w_a -= i_a * (b2_cross(j.r_a, p) + j.impulse3[2]);
        v_b += m_b * p;
        w_b += i_b * (b2_cross(j.r_b, p) + j.impulse3[2]);
      } else {
        j.impulse3[0] = j.impulse3[1] = j.impulse3[2] = 0.0f;
      }
    } else {  // distance: private joints/b2_distance_joint.rs:80-184
      j.u = c_b + j.r_b - c_a - j.r_a;
      j.current_length = j.u.length();
      if (j.current_length > LINEAR_SLOP) {
        j.u *= 1.0f / j.current_length;
      } else {
        j.u.set(0.0f, 0.0f);
        j.mass = 0.0f;
        j.impulse = 0.0f;
        j.lower_impulse = 0.0f;
        j.upper_impulse = 0.0f;
      }
      float cr_au = b2_cross(j.r_a, j.u), cr_bu = b2_cross(j.r_b, j.u);
      float inv_mass = j.inv_mass_a + j.inv_ia * cr_au * cr_au + j.inv_mass_b + j.inv_ib * cr_bu * cr_bu;
      j.mass = inv_mass != 0.0f ? 1.0f / inv_mass : 0.0f;
      if (j.stiffness > 0.0f && j.min_length < j.max_length) {  // soft
        float c = j.current_length - j.length;
        float d = j.damping, k = j.stiffness, h = step.dt;
        j.gamma = h * (d + h * k);
        j.gamma = j.gamma != 0.0f ? 1.0f / j.gamma : 0.0f;
        j.bias = c * h * k * j.gamma;
        inv_mass += j.gamma;
        j.soft_mass = inv_mass != 0.0f ? 1.0f / inv_mass : 0.0f;
      } else {  // rigid
        j.gamma = 0.0f;
        j.bias = 0.0f;
        j.soft_mass = j.mass;
      }
      if (step.warm_starting) {
        j.impulse *= step.dt_ratio;
        j.lower_impulse *= step.dt_ratio;
        j.upper_impulse *= step.dt_ratio;
        Vec2 p = (j.impulse + j.lower_impulse - j.upper_impulse) * j.u;
        v_a -= j.inv_mass_a * p;
        w_a -= j.inv_ia * b2_cross(j.r_a, p);
        v_b += j.inv_mass_b * p;
        w_b += j.inv_ib * b2_cross(j.r_b, p);
      } else {
        j.impulse = 0.0f;
      }
    }
    is.velocities[j.index_a].v = v_a;
    is.velocities[j.index_a].w = w_a;
    is.velocities[j.index_b].v = v_b;
    is.velocities[j.index_b].w = w_b;
  }
  void joint_solve_velocity_constraints(Joint& j, const TimeStep& step, Island& is) {
    if (j.type == J_GEAR) { gear_solve_velocity(j, is); return; }
    Vec2 v_a = is.velocities[j.index_a].v;
    float w_a = is.velocities[j.index_a].w;
    Vec2 v_b = is.velocities[j.index_b].v;
    float w_b = is.velocities[j.index_b].w;
    if (j.type == J_REVOLUTE) {  // private joints/b2_revolute_joint.rs:125-217
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      bool fixed_rotation = i_a + i_b == 0.0f;
      if (j.enable_motor && fixed_rotation == false) {
        float cdot = w_b - w_a - j.motor_speed;
        float impulse = -j.axial_mass * cdot;
        float old_impulse = j.motor_impulse;
        float max_impulse = step.dt * j.max_motor_torque;
        j.motor_impulse = b2_clamp(j.motor_impulse + impulse, -max_impulse, max_impulse);
        impulse = j.motor_impulse - old_impulse;
        w_a -= i_a * impulse;
        w_b += i_b * impulse;
      }
      if (j.enable_limit && fixed_rotation == false) {
        {  // lower limit
          float c = j.angle - j.lower_angle;
          float cdot = w_b - w_a;
          float impulse = -j.axial_mass * (cdot + b2_max(c, 0.0f) * step.inv_dt);
          float old_impulse = j.lower_impulse;
          j.lower_impulse = b2_max(j.lower_impulse + impulse, 0.0f);
          impulse = j.lower_impulse - old_impulse;
          w_a -= i_a * impulse;
          w_b += i_b * impulse;
        }
        {  // upper limit (signs flipped)
          float c = j.upper_angle - j.angle;
          float cdot = w_a - w_b;
          float impulse = -j.axial_mass * (cdot + b2_max(c, 0.0f) * step.inv_dt);
          float old_impulse = j.upper_impulse;
          j.upper_impulse = b2_max(j.upper_impulse + impulse, 0.0f);
          impulse = j.upper_impulse - old_impulse;
          w_a += i_a * impulse;
          w_b -= i_b * impulse;
        }
      }
      {  // point-to-point constraint
        Vec2 cdot = v_b + b2_cross_sv(w_b, j.r_b) - v_a - b2_cross_sv(w_a, j.r_a);
        Vec2 impulse = j.k.solve(-cdot);
        j.impulse2.x += impulse.x;
        j.impulse2.y += impulse.y;
        v_a -= m_a * impulse;
        w_a -= i_a * b2_cross(j.r_a, impulse);
        v_b += m_b * impulse;
        w_b += i_b * b2_cross(j.r_b, impulse);
      }
    } else if (j.type == J_PRISMATIC) {  // private joints/b2_prismatic_joint.rs:282-393
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      if (j.enable_motor) {  // linear motor
        float cdot = b2_dot(j.axis, v_b - v_a) + j.a2 * w_b - j.a1 * w_a;
        float impulse = j.axial_mass * (j.motor_speed - cdot);
        float old_impulse = j.motor_impulse;
        float max_impulse = step.dt * j.max_motor_torque;
        j.motor_impulse = b2_clamp(j.motor_impulse + impulse, -max_impulse, max_impulse);
        impulse = j.motor_impulse - old_impulse;
        Vec2 p = impulse * j.axis;
        float la = impulse * j.a1, lb = impulse * j.a2;
        v_a -= m_a * p;
        w_a -= i_a * la;
        v_b += m_b * p;
        w_b += i_b * lb;
      }
      if (j.enable_limit) {
        {  // lower limit
          float c = j.translation - j.lower_angle;
          float cdot = b2_dot(j.axis, v_b - v_a) + j.a2 * w_b - j.a1 * w_a;
          float impulse = -j.axial_mass * (cdot + b2_max(c, 0.0f) * step.inv_dt);
          float old_impulse = j.lower_impulse;
          j.lower_impulse = b2_max(j.lower_impulse + impulse, 0.0f);
          impulse = j.lower_impulse - old_impulse;
          Vec2 p = impulse * j.axis;
          float la = impulse * j.a1, lb = impulse * j.a2;
          v_a -= m_a * p;
          w_a -= i_a * la;
          v_b += m_b * p;
          w_b += i_b * lb;
        }
        {  // upper limit: signs flipped to keep c positive when the constraint is satisfied
          float c = j.upper_angle - j.translation;
          float cdot = b2_dot(j.axis, v_a - v_b) + j.a1 * w_a - j.a2 * w_b;
          float impulse = -j.axial_mass * (cdot + b2_max(c, 0.0f) * step.inv_dt);
          float old_impulse = j.upper_impulse;
          j.upper_impulse = b2_max(j.upper_impulse + impulse, 0.0f);
          impulse = j.upper_impulse - old_impulse;
          Vec2 p = impulse * j.axis;
          float la = impulse * j.a1, lb = impulse * j.a2;
          v_a += m_a * p;
          w_a += i_a * la;
          v_b -= m_b * p;
          w_b -= i_b * lb;
        }
      }
      {  // prismatic constraint in 2D
        Vec2 cdot(b2_dot(j.perp, v_b - v_a) + j.s2 * w_b - j.s1 * w_a, w_b - w_a);
        Vec2 df = j.k.solve(-cdot);
        j.impulse2 += df;
        Vec2 p = df.x * j.perp;
        float la = df.x * j.s1 + df.y;
        float lb = df.x * j.s2 + df.y;
        v_a -= m_a * p;
        w_a -= i_a * la;
        v_b += m_b * p;
        w_b += i_b * lb;
      }
    } else if (j.type == J_PULLEY) {  // private joints/b2_pulley_joint.rs:139-170
      Vec2 vp_a = v_a + b2_cross_sv(w_a, j.r_a);
      Vec2 vp_b = v_b + b2_cross_sv(w_b, j.r_b);
      float cdot = -b2_dot(j.u, vp_a) - j.ratio * b2_dot(j.u_b, vp_b);
      float impulse = -j.mass * cdot;
      j.impulse += impulse;
      Vec2 pa = -impulse * j.u;
      Vec2 pb = -j.ratio * impulse * j.u_b;
      v_a += j.inv_mass_a * pa;
      w_a += j.inv_ia * b2_cross(j.r_a, pa);
      v_b += j.inv_mass_b * pb;
      w_b += j.inv_ib * b2_cross(j.r_b, pb);
    } else if (j.type == J_MOUSE) {  // private joints/b2_mouse_joint.rs:69-92
      Vec2 cdot = v_b + b2_cross_sv(w_b, j.r_b);
      Vec2 t = -(cdot + j.linear_error + j.gamma * j.impulse2);
      Vec2 impulse(j.k.ex.x * t.x + j.k.ey.x * t.y, j.k.ex.y * t.x + j.k.ey.y * t.y);  // b2_mul(Mat22, v)
      Vec2 old_impulse = j.impulse2;
      j.impulse2 += impulse;
      float max_impulse = step.dt * j.max_force;
      if (j.impulse2.length_squared() > max_impulse * max_impulse) j.impulse2 *= max_impulse / j.impulse2.length();
      impulse = j.impulse2 - old_impulse;
      v_b += j.inv_mass_b * impulse;
      w_b += j.inv_ib * b2_cross(j.r_b, impulse);
    } else if (j.type == J_FRICTION || j.type == J_MOTOR) {  // b2_friction_joint.rs:74-128, b2_motor_joint.rs:98-160
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      float h = step.dt, inv_h = step.inv_dt;
      {  // angular
        float cdot = w_b - w_a;
        if (j.type == J_MOTOR) cdot = w_b - w_a + inv_h * j.correction_factor * j.angular_error;
        float impulse = -j.axial_mass * cdot;
        float old_impulse = j.motor_impulse;
        float max_impulse = h * j.max_motor_torque;
        j.motor_impulse = b2_clamp(j.motor_impulse + impulse, -max_impulse, max_impulse);
        impulse = j.motor_impulse - old_impulse;
        w_a -= i_a * impulse;
        w_b += i_b * impulse;
      }
      {  // linear
        Vec2 cdot = v_b + b2_cross_sv(w_b, j.r_b) - v_a - b2_cross_sv(w_a, j.r_a);
        if (j.type == J_MOTOR) cdot = cdot + inv_h * j.correction_factor * j.linear_error;
        Vec2 impulse = -Vec2(j.k.ex.x * cdot.x + j.k.ey.x * cdot.y, j.k.ex.y * cdot.x + j.k.ey.y * cdot.y);  // b2_mul(Mat22, v)
        Vec2 old_impulse = j.impulse2;
        j.impulse2 += impulse;
        float max_impulse = h * j.max_force;
        if (j.impulse2.length_squared() > max_impulse * max_impulse) {
          j.impulse2.normalize();
          j.impulse2 *= max_impulse;
        }
        impulse = j.impulse2 - old_impulse;
        v_a -= m_a * impulse;
        w_a -= i_a * b2_cross(j.r_a, impulse);
        v_b += m_b * impulse;
        w_b += i_b * b2_cross(j.r_b, impulse);
      }
    } else if (j.type == J_WHEEL) {  // private joints/b2_wheel_joint.rs:172-282
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      {  // spring constraint
        float cdot = b2_dot(j.ax, v_b - v_a) + j.s_bx * w_b - j.s_ax * w_a;
        float impulse = -j.spring_mass * (cdot + j.bias + j.gamma * j.spring_impulse);
        j.spring_impulse += impulse;
        Vec2 p = impulse * j.ax;
        float la = impulse * j.s_ax, lb = impulse * j.s_bx;
        v_a -= m_a * p;
        w_a -= i_a * la;
        v_b += m_b * p;
        w_b += i_b * lb;
      }
      {  // rotational motor constraint (runs with motor_mass = 0 when the motor is off)
        float cdot = w_b - w_a - j.motor_speed;
        float impulse = -j.motor_mass * cdot;
        float old_impulse = j.motor_impulse;
        float max_impulse = step.dt * j.max_motor_torque;
        j.motor_impulse = b2_clamp(j.motor_impulse + impulse, -max_impulse, max_impulse);
        impulse = j.motor_impulse - old_impulse;
        w_a -= i_a * impulse;
        w_b += i_b * impulse;
      }
      if (j.enable_limit) {
        {  // lower limit
          float c = j.translation - j.lower_angle;
          float cdot = b2_dot(j.ax, v_b - v_a) + j.s_bx * w_b - j.s_ax * w_a;
          float impulse = -j.axial_mass * (cdot + b2_max(c, 0.0f) * step.inv_dt);
          float old_impulse = j.lower_impulse;
          j.lower_impulse = b2_max(j.lower_impulse + impulse, 0.0f);
          impulse = j.lower_impulse - old_impulse;
          Vec2 p = impulse * j.ax;
          float la = impulse * j.s_ax, lb = impulse * j.s_bx;
          v_a -= m_a * p;
          w_a -= i_a * la;
          v_b += m_b * p;
          w_b += i_b * lb;
        }
        {  // upper limit
          float c = j.upper_angle - j.translation;
          float cdot = b2_dot(j.ax, v_a - v_b) + j.s_ax * w_a - j.s_bx * w_b;
          float impulse = -j.axial_mass * (cdot + b2_max(c, 0.0f) * step.inv_dt);
          float old_impulse = j.upper_impulse;
          j.upper_impulse = b2_max(j.upper_impulse + impulse, 0.0f);
          impulse = j.upper_impulse - old_impulse;
          Vec2 p = impulse * j.ax;
          float la = impulse * j.s_ax, lb = impulse * j.s_bx;
          v_a += m_a * p;
          w_a += i_a * la;
          v_b -= m_b * p;
          w_b -= i_b * lb;
        }
      }
      {  // point to line constraint
        float cdot = b2_dot(j.ay, v_b - v_a) + j.s_by * w_b - j.s_ay * w_a;
        float impulse = -j.mass * cdot;
        j.impulse += impulse;
        Vec2 p = impulse * j.ay;
        float la = impulse * j.s_ay, lb = impulse * j.s_by;
        v_a -= m_a * p;
        w_a -= i_a * la;
        v_b += m_b * p;
        w_b += i_b * lb;
      }
    } else if (j.type == J_WELD) {  // private joints/b2_weld_joint.rs:138-205
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      const float* m = j.m33;
      if (j.stiffness > 0.0f) {
        float cdot2 = w_b - w_a;
        float impulse2 = -m[8] * (cdot2 + j.bias + j.gamma * j.impulse3[2]);
        j.impulse3[2] += impulse2;
        w_a -= i_a * impulse2;
        w_b += i_b * impulse2;
        Vec2 cdot1 = v_b + b2_cross_sv(w_b, j.r_b) - v_a - b2_cross_sv(w_a, j.r_a);
        Vec2 impulse1 = -Vec2(m[0] * cdot1.x + m[3] * cdot1.y, m[1] * cdot1.x + m[4] * cdot1.y);  // b2_mul22
        j.impulse3[0] += impulse1.x;
        j.impulse3[1] += impulse1.y;
        Vec2 p = impulse1;
        v_a -= m_a * p;
        w_a -= i_a * b2_cross(j.r_a, p);
        v_b += m_b * p;
        w_b += i_b * b2_cross(j.r_b, p);
      } else {
        Vec2 cdot1 = v_b + b2_cross_sv(w_b, j.r_b) - v_a - b2_cross_sv(w_a, j.r_a);
        float cdot2 = w_b - w_a;
        // b2_mul_mat33 (src/b2_math.rs:601-603): v.x * ex + v.y * ey + v.z * ez, then the negation
        float ix = -((cdot1.x * m[0] + cdot1.y * m[3]) + cdot2 * m[6]);
        float iy = -((cdot1.x * m[1] + cdot1.y * m[4]) + cdot2 * m[7]);
        float iz = -((cdot1.x * m[2] + cdot1.y * m[5]) + cdot2 * m[8]);
        j.impulse3[0] += ix; j.impulse3[1] += iy; j.impulse3[2] += iz;
        Vec2 p(ix, iy);
        v_a -= m_a * p;
        w_a -= i_a * (b2_cross(j.r_a, p) + iz);
        v_b += m_b * p;
        w_b += i_b * (b2_cross(j.r_b, p) + iz);
      }
    } else {  // distance: private joints/b2_distance_joint.rs:186-277
      if (j.min_length < j.max_length) {
        if (j.stiffness > 0.0f) {
          Vec2 vp_a = v_a + b2_cross_sv(w_a, j.r_a);
          Vec2 vp_b = v_b + b2_cross_sv(w_b, j.r_b);
          float cdot = b2_dot(j.u, vp_b - vp_a);
          float impulse = -j.soft_mass * (cdot + j.bias + j.gamma * j.impulse);
          j.impulse += impulse;
          Vec2 p = impulse * j.u;
          v_a -= j.inv_mass_a * p;
          w_a -= j.inv_ia * b2_cross(j.r_a, p);
          v_b += j.inv_mass_b * p;
          w_b += j.inv_ib * b2_cross(j.r_b, p);
        }
        {  // lower
          float c = j.current_length - j.min_length;
          float bias = b2_max(0.0f, c) * step.inv_dt;
          Vec2 vp_a = v_a + b2_cross_sv(w_a, j.r_a);
          Vec2 vp_b = v_b + b2_cross_sv(w_b, j.r_b);
          float cdot = b2_dot(j.u, vp_b - vp_a);
          float impulse = -j.mass * (cdot + bias);
          float old_impulse = j.lower_impulse;
          j.lower_impulse = b2_max(0.0f, j.lower_impulse + impulse);
          impulse = j.lower_impulse - old_impulse;
          Vec2 p = impulse * j.u;
          v_a -= j.inv_mass_a * p;
          w_a -= j.inv_ia * b2_cross(j.r_a, p);
          v_b += j.inv_mass_b * p;
          w_b += j.inv_ib * b2_cross(j.r_b, p);
        }
        {  // upper
          float c = j.max_length - j.current_length;
          float bias = b2_max(0.0f, c) * step.inv_dt;
          Vec2 vp_a = v_a + b2_cross_sv(w_a, j.r_a);
          Vec2 vp_b = v_b + b2_cross_sv(w_b, j.r_b);
          float cdot = b2_dot(j.u, vp_a - vp_b);
          float impulse = -j.mass * (cdot + bias);
          float old_impulse = j.upper_impulse;
          j.upper_impulse = b2_max(0.0f, j.upper_impulse + impulse);
          impulse = j.upper_impulse - old_impulse;
          Vec2 p = -impulse * j.u;
          v_a -= j.inv_mass_a * p;
          w_a -= j.inv_ia * b2_cross(j.r_a, p);
          v_b += j.inv_mass_b * p;
          w_b += j.inv_ib * b2_cross(j.r_b, p);
        }
      } else {  // equal limits
        Vec2 vp_a = v_a + b2_cross_sv(w_a, j.r_a);
        Vec2 vp_b = v_b + b2_cross_sv(w_b, j.r_b);
        float cdot = b2_dot(j.u, vp_b - vp_a);
        float impulse = -j.mass * cdot;
        j.impulse += impulse;
        Vec2 p = impulse * j.u;
        v_a -= j.inv_mass_a * p;
        w_a -= j.inv_ia * b2_cross(j.r_a, p);
        v_b += j.inv_mass_b * p;
        w_b += j.inv_ib * b2_cross(j.r_b, p);
      }
    }
    is.velocities[j.index_a].v = v_a;
    is.velocities[j.index_a].w = w_a;
    is.velocities[j.index_b].v = v_b;
    is.velocities[j.index_b].w = w_b;
  }
  bool joint_solve_position_constraints(Joint& j, Island& is) {
    if (j.type == J_GEAR) return gear_solve_position(j, is);
    Vec2 c_a = is.positions[j.index_a].c;
    float a_a = is.positions[j.index_a].a;
    Vec2 c_b = is.positions[j.index_b].c;
    float a_b = is.positions[j.index_b].a;
    bool okay;
    if (j.type == J_REVOLUTE) {  // private joints/b2_revolute_joint.rs:219-301
      Rot q_a(a_a), q_b(a_b);
      float angular_error = 0.0f, position_error;
      bool fixed_rotation = j.inv_ia + j.inv_ib == 0.0f;
      if (j.enable_limit && fixed_rotation == false) {
        float angle = a_b - a_a - j.reference_angle;
        float c = 0.0f;
        if (fabsf(j.upper_angle - j.lower_angle) < 2.0f * ANGULAR_SLOP) {
          c = b2_clamp(angle - j.lower_angle, -MAX_ANGULAR_CORRECTION, MAX_ANGULAR_CORRECTION);
        } else if (angle <= j.lower_angle) {
          c = b2_clamp(angle - j.lower_angle + ANGULAR_SLOP, -MAX_ANGULAR_CORRECTION, 0.0f);
        } else if (angle >= j.upper_angle) {
          c = b2_clamp(angle - j.upper_angle - ANGULAR_SLOP, 0.0f, MAX_ANGULAR_CORRECTION);
        }
        float limit_impulse = -j.axial_mass * c;
        a_a -= j.inv_ia * limit_impulse;
        a_b += j.inv_ib * limit_impulse;
        angular_error = fabsf(c);
      }
      {
        q_a.set(a_a);
        q_b.set(a_b);
        Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
        Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
        Vec2 c = c_b + r_b - c_a - r_a;
        position_error = c.length();
        float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
        Mat22 k;
        k.ex.x = m_a + m_b + i_a * r_a.y * r_a.y + i_b * r_b.y * r_b.y;
        k.ex.y = -i_a * r_a.x * r_a.y - i_b * r_b.x * r_b.y;
        k.ey.x = k.ex.y;
        k.ey.y = m_a + m_b + i_a * r_a.x * r_a.x + i_b * r_b.x * r_b.x;
        Vec2 impulse = -k.solve(c);
        c_a -= m_a * impulse;
        a_a -= i_a * b2_cross(r_a, impulse);
        c_b += m_b * impulse;
        a_b += i_b * b2_cross(r_b, impulse);
      }
      okay = position_error <= LINEAR_SLOP && angular_error <= ANGULAR_SLOP;
    } else if (j.type == J_PRISMATIC) {  // private joints/b2_prismatic_joint.rs:395-506
      Rot q_a(a_a), q_b(a_b);
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
      Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
      Vec2 d = c_b + r_b - c_a - r_a;
      Vec2 axis = b2_mul_rot(q_a, j.local_xaxis_a);
      float a1 = b2_cross(d + r_a, axis);
      float a2 = b2_cross(r_b, axis);
      Vec2 perp = b2_mul_rot(q_a, j.local_yaxis_a);
      float s1 = b2_cross(d + r_a, perp);
      float s2 = b2_cross(r_b, perp);
      float imp[3];
      Vec2 c1(b2_dot(perp, d), a_b - a_a - j.reference_angle);
      float linear_error = fabsf(c1.x);
      float angular_error = fabsf(c1.y);
      bool active = false;
      float c2 = 0.0f;
      if (j.enable_limit) {
        float translation = b2_dot(axis, d);
        if (fabsf(j.upper_angle - j.lower_angle) < 2.0f * LINEAR_SLOP) {
          c2 = translation;
          linear_error = b2_max(linear_error, fabsf(translation));
          active = true;
        } else if (translation <= j.lower_angle) {
          c2 = b2_min(translation - j.lower_angle, 0.0f);
          linear_error = b2_max(linear_error, j.lower_angle - translation);
          active = true;
        } else if (translation >= j.upper_angle) {
          c2 = b2_max(translation - j.upper_angle, 0.0f);
          linear_error = b2_max(linear_error, translation - j.upper_angle);
          active = true;
        }
      }
      if (active) {
        float k11 = m_a + m_b + i_a * s1 * s1 + i_b * s2 * s2;
        float k12 = i_a * s1 + i_b * s2;
        float k13 = i_a * s1 * a1 + i_b * s2 * a2;
        float k22 = i_a + i_b;
        if (k22 == 0.0f) k22 = 1.0f;  // fixed rotation
        float k23 = i_a * a1 + i_b * a2;
        float k33 = m_a + m_b + i_a * a1 * a1 + i_b * a2 * a2;
        float k[9] = {k11, k12, k13, k12, k22, k23, k13, k23, k33};
        float c[3] = {-c1.x, -c1.y, -c2};
        mat33_solve33(k, c, imp);
      } else {
        float k11 = m_a + m_b + i_a * s1 * s1 + i_b * s2 * s2;
        float k12 = i_a * s1 + i_b * s2;
        float k22 = i_a + i_b;
        if (k22 == 0.0f) k22 = 1.0f;
        Mat22 k;
        k.ex.set(k11, k12);
        k.ey.set(k12, k22);
        Vec2 impulse1 = k.solve(-c1);
        imp[0] = impulse1.x; imp[1] = impulse1.y; imp[2] = 0.0f;
      }
      Vec2 p = imp[0] * perp + imp[2] * axis;
      float la = imp[0] * s1 + imp[1] + imp[2] * a1;
      float lb = imp[0] * s2 + imp[1] + imp[2] * a2;
      c_a -= m_a * p;
      a_a -= i_a * la;
      c_b += m_b * p;
      a_b += i_b * lb;
      okay = linear_error <= LINEAR_SLOP && angular_error <= ANGULAR_SLOP;
    } else if (j.type == J_FRICTION || j.type == J_MOTOR || j.type == J_MOUSE) {  // no position rows: always within tolerance
      return true;
    } else if (j.type == J_PULLEY) {  // private joints/b2_pulley_joint.rs:172-240
      Rot q_a(a_a), q_b(a_b);
      Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
      Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
      Vec2 u_a = c_a + r_a - j.ground_anchor_a;
      Vec2 u_b = c_b + r_b - j.ground_anchor_b;
      float length_a = u_a.length(), length_b = u_b.length();
      if (length_a > 10.0f * LINEAR_SLOP) u_a *= 1.0f / length_a; else u_a.set_zero();
      if (length_b > 10.0f * LINEAR_SLOP) u_b *= 1.0f / length_b; else u_b.set_zero();
      float ru_a = b2_cross(r_a, u_a), ru_b = b2_cross(r_b, u_b);
      float m_a = j.inv_mass_a + j.inv_ia * ru_a * ru_a;
      float m_b = j.inv_mass_b + j.inv_ib * ru_b * ru_b;
      float mass = m_a + j.ratio * j.ratio * m_b;
      if (mass > 0.0f) mass = 1.0f / mass;
      float c = j.constant - length_a - j.ratio * length_b;
      float linear_error = fabsf(c);
      float impulse = -mass * c;
      Vec2 pa = -impulse * u_a;
      Vec2 pb = -j.ratio * impulse * u_b;
      c_a += j.inv_mass_a * pa;
      a_a += j.inv_ia * b2_cross(r_a, pa);
      c_b += j.inv_mass_b * pb;
      a_b += j.inv_ib * b2_cross(r_b, pb);
      okay = linear_error < LINEAR_SLOP;
    } else if (j.type == J_WHEEL) {  // private joints/b2_wheel_joint.rs:284-380
      float linear_error = 0.0f;
      if (j.enable_limit) {
        Rot q_a(a_a), q_b(a_b);
        Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
        Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
        Vec2 d = (c_b - c_a) + r_b - r_a;
        Vec2 ax = b2_mul_rot(q_a, j.local_xaxis_a);
        float s_ax = b2_cross(d + r_a, j.ax);  // the reference crosses with m_ax of init_velocity_constraints here
        float s_bx = b2_cross(r_b, j.ax);
        float c = 0.0f;
        float translation = b2_dot(ax, d);
        if (fabsf(j.upper_angle - j.lower_angle) < 2.0f * LINEAR_SLOP) c = translation;
        else if (translation <= j.lower_angle) c = b2_min(translation - j.lower_angle, 0.0f);
        else if (translation >= j.upper_angle) c = b2_max(translation - j.upper_angle, 0.0f);
        if (c != 0.0f) {
          float inv_mass = j.inv_mass_a + j.inv_mass_b + j.inv_ia * s_ax * s_ax + j.inv_ib * s_bx * s_bx;
          float impulse = 0.0f;
          if (inv_mass != 0.0f) impulse = -c / inv_mass;
          Vec2 p = impulse * ax;
          float la = impulse * s_ax, lb = impulse * s_bx;
          c_a -= j.inv_mass_a * p;
          a_a -= j.inv_ia * la;
          c_b += j.inv_mass_b * p;
          a_b += j.inv_ib * lb;
          linear_error = fabsf(c);
        }
      }
      {  // solve perpendicular constraint
        Rot q_a(a_a), q_b(a_b);
        Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
        Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
        Vec2 d = (c_b - c_a) + r_b - r_a;
        Vec2 ay = b2_mul_rot(q_a, j.local_yaxis_a);
        float s_ay = b2_cross(d + r_a, ay);
        float s_by = b2_cross(r_b, ay);
        float c = b2_dot(d, ay);
        // the reference uses m_s_ay / m_s_by of init_velocity_constraints in the effective mass
        float inv_mass = j.inv_mass_a + j.inv_mass_b + j.inv_ia * j.s_ay * j.s_ay + j.inv_ib * j.s_by * j.s_by;
        float impulse = 0.0f;
        if (inv_mass != 0.0f) impulse = -c / inv_mass;
        Vec2 p = impulse * ay;
        float la = impulse * s_ay, lb = impulse * s_by;
        c_a -= j.inv_mass_a * p;
        a_a -= j.inv_ia * la;
        c_b += j.inv_mass_b * p;
        a_b += j.inv_ib * lb;
        linear_error = b2_max(linear_error, fabsf(c));
      }
      okay = linear_error <= LINEAR_SLOP;
    } else if (j.type == J_WELD) {  // private joints/b2_weld_joint.rs:207-283
      Rot q_a(a_a), q_b(a_b);
      float m_a = j.inv_mass_a, m_b = j.inv_mass_b, i_a = j.inv_ia, i_b = j.inv_ib;
      Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
      Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
      float position_error, angular_error;
      float k[9];
      k[0] = m_a + m_b + r_a.y * r_a.y * i_a + r_b.y * r_b.y * i_b;
      k[3] = -r_a.y * r_a.x * i_a - r_b.y * r_b.x * i_b;
      k[6] = -r_a.y * i_a - r_b.y * i_b;
      k[1] = k[3];
      k[4] = m_a + m_b + r_a.x * r_a.x * i_a + r_b.x * r_b.x * i_b;
      k[7] = r_a.x * i_a + r_b.x * i_b;
      k[2] = k[6];
      k[5] = k[7];
      k[8] = i_a + i_b;
      auto solve22 = [&](Vec2 b) {  // B2Mat33::solve22 (private b2_math.rs:19-30)
        float a11 = k[0], a12 = k[3], a21 = k[1], a22 = k[4];
        float det = a11 * a22 - a12 * a21;
        if (det != 0.0f) det = 1.0f / det;
        return Vec2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
      };
      if (j.stiffness > 0.0f) {
        Vec2 c1 = c_b + r_b - c_a - r_a;
        position_error = c1.length();
        angular_error = 0.0f;
        Vec2 p = -solve22(c1);
        c_a -= m_a * p;
        a_a -= i_a * b2_cross(r_a, p);
        c_b += m_b * p;
        a_b += i_b * b2_cross(r_b, p);
      } else {
        Vec2 c1 = c_b + r_b - c_a - r_a;
        float c2 = a_b - a_a - j.reference_angle;
        position_error = c1.length();
        angular_error = fabsf(c2);
        float imp[3];
        if (k[8] > 0.0f) {
          float c[3] = {c1.x, c1.y, c2}, x[3];
          mat33_solve33(k, c, x);
          imp[0] = -x[0]; imp[1] = -x[1]; imp[2] = -x[2];
        } else {
          Vec2 impulse2 = -solve22(c1);
          imp[0] = impulse2.x; imp[1] = impulse2.y; imp[2] = 0.0f;
        }
        Vec2 p(imp[0], imp[1]);
        c_a -= m_a * p;
        a_a -= i_a * (b2_cross(r_a, p) + imp[2]);
        c_b += m_b * p;
        a_b += i_b * (b2_cross(r_b, p) + imp[2]);
      }
      okay = position_error <= LINEAR_SLOP && angular_error <= ANGULAR_SLOP;
    } else {  // distance: private joints/b2_distance_joint.rs:279-320
      Rot q_a(a_a), q_b(a_b);
      Vec2 r_a = b2_mul_rot(q_a, j.local_anchor_a - j.local_center_a);
      Vec2 r_b = b2_mul_rot(q_b, j.local_anchor_b - j.local_center_b);
      Vec2 u = c_b + r_b - c_a - r_a;
      float length = u.normalize();
      float c;
      if (j.min_length == j.max_length) c = length - j.min_length;
      else if (length < j.min_length) c = length - j.min_length;
      else if (j.max_length < length) c = length - j.max_length;
      else return true;  // positions untouched
      float impulse = -j.mass * c;
      Vec2 p = impulse * u;
      c_a -= j.inv_mass_a * p;
      a_a -= j.inv_ia * b2_cross(r_a, p);
      c_b += j.inv_mass_b * p;
      a_b += j.inv_ib * b2_cross(r_b, p);
      okay = fabsf(c) < LINEAR_SLOP;
    }
    is.positions[j.index_a].c = c_a;
    is.positions[j.index_a].a = a_a;
    is.positions[j.index_b].c = c_b;
    is.positions[j.index_b].a = a_b;
    return okay;
  }

  // Diagnostic (not part of the reference): depth of the dependency DAG of one in-order
  // sweep over the island's constraints, where only bodies with non-zero inverse mass or
  // inertia create dependencies.
  int wavefront_depth(const Island& is, int sweeps) const {
    std::vector<int> last(is.bodies.size(), 0);
    int depth = 0;
    for (int s = 0; s < sweeps; ++s)
      for (auto& vc : vcs) {
        bool dyn_a = vc.inv_mass_a != 0.0f || vc.inv_ia != 0.0f;
        bool dyn_b = vc.inv_mass_b != 0.0f || vc.inv_ib != 0.0f;
        int l = 0;
        if (dyn_a) l = std::max(l, last[vc.index_a]);
        if (dyn_b) l = std::max(l, last[vc.index_b]);
        ++l;
        if (dyn_a) last[vc.index_a] = l;
        if (dyn_b) last[vc.index_b] = l;
        depth = std::max(depth, l);
      }
    return depth;
  }

  // Diagnostic (not part of the reference): the same DAG for the largest island of the step, plus a simulation of
  // the chunked dataflow schedule of the GPU's giant-island solver (b2g_large.h: the island's constraint list is cut
  // into `workers` contiguous chunks, each worker walks its chunk in order sweep after sweep and waits for the
  // previous visit of either body; one visit = 1 time unit, a hand-over between workers costs `handover` units).
  struct DagStats {
    int contacts = 0, bodies = 0, sweeps = 0, depth = 0, depth_one_sweep = 0;
    double makespan[4] = {0, 0, 0, 0};  // workers = 256, 1024, 4096, 16384
    // cyclic schedule: chunks of `chunk` consecutive constraints dealt round-robin to `workers` threads (the dataflow
    // solver's mapping): makespan_cyclic[i][j] for chunk = {1, 4, 16, 64}[i], workers = {2048, 16384, 65536}[j]
    double makespan_cyclic[4][3] = {{0}};
  };
  DagStats dag;
  void dag_collect(const Island& is, int sweeps, double handover) {
    if ((int)vcs.size() <= dag.contacts) return;
    dag = DagStats();
    dag.contacts = (int)vcs.size();
    dag.bodies = (int)is.bodies.size();
    dag.sweeps = sweeps;
    dag.depth = wavefront_depth(is, sweeps);
    dag.depth_one_sweep = wavefront_depth(is, 1);
    const int workers[4] = {256, 1024, 4096, 16384};
    for (int wi = 0; wi < 4; ++wi) {
      const int P = workers[wi];
      const int chunk = ((int)vcs.size() + P - 1) / P;
      std::vector<double> bfin(is.bodies.size(), 0.0), tfin(P, 0.0);
      std::vector<int> bown(is.bodies.size(), -1);
      double end = 0.0;
      for (int s = 0; s < sweeps; ++s)
        for (size_t k = 0; k < vcs.size(); ++k) {
          const ContactVelocityConstraint& vc = vcs[k];
          const int p = (int)k / chunk;
          const bool dyn_a = vc.inv_mass_a != 0.0f || vc.inv_ia != 0.0f;
          const bool dyn_b = vc.inv_mass_b != 0.0f || vc.inv_ib != 0.0f;
          double start = tfin[p];
          if (dyn_a) start = std::max(start, bfin[vc.index_a] + (bown[vc.index_a] != p && bown[vc.index_a] >= 0 ? handover : 0.0));
          if (dyn_b) start = std::max(start, bfin[vc.index_b] + (bown[vc.index_b] != p && bown[vc.index_b] >= 0 ? handover : 0.0));
          const double fin = start + 1.0;
          tfin[p] = fin;
          if (dyn_a) { bfin[vc.index_a] = fin; bown[vc.index_a] = p; }
          if (dyn_b) { bfin[vc.index_b] = fin; bown[vc.index_b] = p; }
          end = std::max(end, fin);
        }
      dag.makespan[wi] = end;
    }
    const int chunks[4] = {1, 4, 16, 64}, nworkers[3] = {2048, 16384, 65536};
    for (int ci = 0; ci < 4; ++ci)
      for (int wi = 0; wi < 3; ++wi) {
        const int c = chunks[ci], P = nworkers[wi];
        std::vector<double> bfin(is.bodies.size(), 0.0), tfin(P, 0.0);
        std::vector<int> bown(is.bodies.size(), -1);
        double end = 0.0;
        for (int s = 0; s < sweeps; ++s)
          for (size_t k = 0; k < vcs.size(); ++k) {
            const ContactVelocityConstraint& vc = vcs[k];
            const int p = (int)((k / c) % P);
            const bool dyn_a = vc.inv_mass_a != 0.0f || vc.inv_ia != 0.0f;
            const bool dyn_b = vc.inv_mass_b != 0.0f || vc.inv_ib != 0.0f;
            double start = tfin[p];
            if (dyn_a) start = std::max(start, bfin[vc.index_a] + (bown[vc.index_a] != p && bown[vc.index_a] >= 0 ? handover : 0.0));
            if (dyn_b) start = std::max(start, bfin[vc.index_b] + (bown[vc.index_b] != p && bown[vc.index_b] >= 0 ? handover : 0.0));
            const double fin = start + 1.0;
            tfin[p] = fin;
            if (dyn_a) { bfin[vc.index_a] = fin; bown[vc.index_a] = p; }
            if (dyn_b) { bfin[vc.index_b] = fin; bown[vc.index_b] = p; }
            end = std::max(end, fin);
          }
        dag.makespan_cyclic[ci][wi] = end;
      }
  }

  void island_solve(Island& is, const TimeStep& step) {  // b2_island_private.rs:129-328
    double t0 = now_ms();
    float h = step.dt;
    is.positions.resize(is.bodies.size());
    is.velocities.resize(is.bodies.size());
    for (size_t i = 0; i < is.bodies.size(); ++i) {
      Body& b = bodies[is.bodies[i]];
      Vec2 c = b.sweep.c;
      float a = b.sweep.a;
      Vec2 v = b.linear_velocity;
      float w = b.angular_velocity;
      b.sweep.c0 = b.sweep.c;
      b.sweep.a0 = b.sweep.a;
      if (b.type == DYNAMIC_BODY) {
        v += (h * b.inv_mass) * ((b.gravity_scale * b.mass) * gravity + b.force);
        w += h * b.inv_i * b.torque;
        v *= 1.0f / (1.0f + h * b.linear_damping);
        w *= 1.0f / (1.0f + h * b.angular_damping);
      }
      is.positions[i].c = c;
      is.positions[i].a = a;
      is.velocities[i].v = v;
      is.velocities[i].w = w;
    }
    solver_new(is, step);
    initialize_velocity_constraints(is);
    if (step.warm_starting) warm_start(is);
    for (int ji : is.joints) joint_init_velocity_constraints(joints[ji], step, is);  // b2_island_private.rs:198-201
    if (collect_levels) stats.solver_levels = std::max(stats.solver_levels, wavefront_depth(is, step.velocity_iterations));
    if (collect_levels && collect_dag) dag_collect(is, step.velocity_iterations + (step.warm_starting ? 1 : 0), dag_handover);
    double t1 = now_ms();
    profile.solve_init += t1 - t0;
    for (int it = 0; it < step.velocity_iterations; ++it) {  // :207-215: joints first, then contacts
      for (int ji : is.joints) joint_solve_velocity_constraints(joints[ji], step, is);
      solve_velocity_constraints(is);
    }
    store_impulses(is);
    double t2 = now_ms();
    profile.solve_velocity += t2 - t1;
    for (size_t i = 0; i < is.bodies.size(); ++i) {
      Vec2 c = is.positions[i].c;
      float a = is.positions[i].a;
      Vec2 v = is.velocities[i].v;
      float w = is.velocities[i].w;
      Vec2 translation = h * v;
      if (b2_dot(translation, translation) > MAX_TRANSLATION_SQUARED) {
        float ratio = MAX_TRANSLATION / translation.length();
        v *= ratio;
      }
      float rotation = h * w;
      if (rotation * rotation > MAX_ROTATION_SQUARED) {
        float ratio = MAX_ROTATION / fabsf(rotation);
        w *= ratio;
      }
      c += h * v;
      a += h * w;
      is.positions[i].c = c;
      is.positions[i].a = a;
      is.velocities[i].v = v;
      is.velocities[i].w = w;
    }
    bool position_solved = false;
    for (int it = 0; it < step.position_iterations; ++it) {
      bool contacts_okay = solve_position_constraints(is);
      bool joints_okay = true;  // :262-266
      for (int ji : is.joints) {
        bool joint_okay = joint_solve_position_constraints(joints[ji], is);
        joints_okay = joints_okay && joint_okay;
      }
      if (contacts_okay && joints_okay) { position_solved = true; break; }
    }
    for (size_t i = 0; i < is.bodies.size(); ++i) {
      Body& body = bodies[is.bodies[i]];
      body.sweep.c = is.positions[i].c;
      body.sweep.a = is.positions[i].a;
      body.linear_velocity = is.velocities[i].v;
      body.angular_velocity = is.velocities[i].w;
      synchronize_transform(body);
    }
    profile.solve_position += now_ms() - t2;
    for (size_t i = 0; i < is.contacts.size(); ++i) {  // self_.report(&contact_solver.m_velocity_constraints)
      const Contact& c = contacts[is.contacts[i]];
      const ContactVelocityConstraint& vc = vcs[i];
      PostSolveEvent e = {c.fixture_a, c.index_a, c.fixture_b, c.index_b, vc.point_count, {0.0f, 0.0f}, {0.0f, 0.0f}};
      for (int j = 0; j < vc.point_count; ++j) {
        e.normal_impulses[j] = vc.points[j].normal_impulse;
        e.tangent_impulses[j] = vc.points[j].tangent_impulse;
      }
      post_solve_events.push_back(e);
    }
    if (allow_sleep) {
      float min_sleep_time = MAX_FLOAT;
      const float lin_tol_sqr = LINEAR_SLEEP_TOLERANCE * LINEAR_SLEEP_TOLERANCE;
      const float ang_tol_sqr = ANGULAR_SLEEP_TOLERANCE * ANGULAR_SLEEP_TOLERANCE;
      for (int bi : is.bodies) {
        Body& b = bodies[bi];
        if (b.type == STATIC_BODY) continue;
        if (!(b.flags & BF_AUTO_SLEEP) || b.angular_velocity * b.angular_velocity > ang_tol_sqr ||
            b2_dot(b.linear_velocity, b.linear_velocity) > lin_tol_sqr) {
          b.sleep_time = 0.0f;
          min_sleep_time = 0.0f;
        } else {
          b.sleep_time += h;
          min_sleep_time = b2_min(min_sleep_time, b.sleep_time);
        }
      }
      if (min_sleep_time >= TIME_TO_SLEEP && position_solved)
        for (int bi : is.bodies) set_awake(bi, false);
    }
  }

  void solve(const TimeStep& step) {  // b2_world.rs(private):356-531
    profile.solve_init = profile.solve_velocity = profile.solve_position = 0.0;
    Island island;
    for (int b = body_list; b != -1; b = bodies[b].next) bodies[b].flags &= ~BF_ISLAND;
    for (int c = contact_list; c != -1; c = contacts[c].next) contacts[c].flags &= ~CF_ISLAND;
    for (Joint& j : joints) j.island_flag = false;
    std::vector<int> stack;
    stack.reserve(bodies.size());
    for (int seed = body_list; seed != -1; seed = bodies[seed].next) {
      {
        const Body& s = bodies[seed];
        if (s.flags & BF_ISLAND) continue;
        if (!(s.flags & BF_AWAKE) || !(s.flags & BF_ENABLED)) continue;
        if (s.type == STATIC_BODY) continue;
      }
      island.clear();
      stack.clear();
      stack.push_back(seed);
      bodies[seed].flags |= BF_ISLAND;
      while (!stack.empty()) {
        int bi = stack.back();
        stack.pop_back();
        bodies[bi].island_index = (int)island.bodies.size();  // b2_island.rs:52-55
        island.bodies.push_back(bi);
        if (bodies[bi].type == STATIC_BODY) continue;
        bodies[bi].flags |= BF_AWAKE;
        for (int e = bodies[bi].contact_list; e != -1; e = edge(e).next) {
          Contact& contact = contacts[e >> 1];
          if (contact.flags & CF_ISLAND) continue;
          if (!(contact.flags & CF_ENABLED) || !(contact.flags & CF_TOUCHING)) continue;
          if (fixtures[contact.fixture_a].is_sensor || fixtures[contact.fixture_b].is_sensor) continue;
          island.contacts.push_back(e >> 1);
          contact.flags |= CF_ISLAND;
          int other = edge(e).other;
          if (bodies[other].flags & BF_ISLAND) continue;
          stack.push_back(other);
          bodies[other].flags |= BF_ISLAND;
        }
        // search all joints connected to this body (b2_world.rs(private):461-483), newest edge first
        const std::vector<int>& je = bodies[bi].joint_edges;
        for (size_t q = je.size(); q-- > 0;) {
          Joint& joint = joints[je[q] >> 1];
          if (joint.island_flag) continue;
          int other = (je[q] & 1) ? joint.body_a : joint.body_b;
          if (!(bodies[other].flags & BF_ENABLED)) continue;  // don't simulate joints connected to disabled bodies
          island.joints.push_back(je[q] >> 1);
          joint.island_flag = true;
          if (bodies[other].flags & BF_ISLAND) continue;
          stack.push_back(other);
          bodies[other].flags |= BF_ISLAND;
        }
      }
      island_solve(island, step);
      ++stats.islands;
      stats.island_bodies += (int)island.bodies.size();
      stats.island_contacts += (int)island.contacts.size();
      for (int bi : island.bodies)
        if (bodies[bi].type == STATIC_BODY) bodies[bi].flags &= ~BF_ISLAND;
    }
    double t0 = now_ms();
    for (int b = body_list; b != -1; b = bodies[b].next) {
      if (!(bodies[b].flags & BF_ISLAND)) continue;
      if (bodies[b].type == STATIC_BODY) continue;
      synchronize_fixtures(b);
    }
    find_new_contacts();
    profile.broadphase = now_ms() - t0;
  }

  void step(float dt, int velocity_iterations, int position_iterations) {  // b2_world.rs(private):903-959
    double ts = now_ms();
    stats = StepStats();
    dag = DagStats();
    events.clear();
    post_solve_events.clear();
    if (new_contacts) {
      find_new_contacts();
      new_contacts = false;
    }
    locked = true;
    TimeStep st;
    st.dt = dt;
    st.velocity_iterations = velocity_iterations;
    st.position_iterations = position_iterations;
    st.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
    st.dt_ratio = inv_dt0 * dt;
    st.warm_starting = warm_starting;
    {
      double t = now_ms();
      collide();
      profile.collide = now_ms() - t;
    }
    for (int c = contact_list; c != -1; c = contacts[c].next)
      if (contacts[c].flags & CF_TOUCHING) ++stats.touching;
    if (step_complete && st.dt > 0.0f) {
      double t = now_ms();
      solve(st);
      profile.solve = now_ms() - t;
    }
    // continuous physics (solve_toi) is out of scope and disabled in both engines
    if (st.dt > 0.0f) inv_dt0 = st.inv_dt;
    if (clear_forces_flag)
      for (int b = body_list; b != -1; b = bodies[b].next) {
        bodies[b].force.set_zero();
        bodies[b].torque = 0.0f;
      }
    locked = false;
    stats.contacts = contact_count;
    for (auto& b : bodies)
      if (b.flags & BF_AWAKE) ++stats.awake_bodies;
    profile.step = now_ms() - ts;
  }
};

}  // namespace b2o
