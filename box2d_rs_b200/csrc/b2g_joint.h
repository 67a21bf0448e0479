// b2g_joint.h — joints inside the island solve: revolute, prismatic, wheel, distance, weld, friction and
// motor joints.
//
// Reference: B2jointTraitDyn::{init_velocity_constraints, solve_velocity_constraints, solve_position_constraints}
// (src/b2_joint.rs:268-286) as driven by B2island::solve (src/private/dynamics/b2_island_private.rs:198-201 init after the
// contact warm start, :207-215 joints BEFORE contacts in every velocity iteration, :257-274 joints AFTER contacts in every
// position iteration, early exit only when both are within tolerance);
//   revolute: src/private/dynamics/joints/b2_revolute_joint.rs:22-123 / :125-217 / :219-301
//   distance: src/private/dynamics/joints/b2_distance_joint.rs:80-184 / :186-277 / :279-320
//   weld:     src/private/dynamics/joints/b2_weld_joint.rs:22-136 / :138-205 / :207-283
//   prismatic: src/private/dynamics/joints/b2_prismatic_joint.rs:168-280 / :282-393 / :395-506 (impulses and switches laid
//              out like the revolute joint's; param 1 / 2 = translation limits, 3 = max motor force, 5 / 6 = local x axis)
// Expression shapes are kept operation for operation (one rounding per operation, no FMA), like the contact solver.
//
// Data: the static part of a joint (type, bodies, anchors, limits, lengths, COLLIDE_CONNECTED) is topology shared by the
// worlds of a batch (Batch::joints); what the solver or the user changes per world lives in two float4 rows
//   j_s0: impulse.x impulse.y motor_impulse lower_impulse     (distance: impulse, -, -, lower_impulse; weld: impulse.x .y .z, -)
//   j_s1: upper_impulse motor_speed max_motor_torque (int bits) ENABLE_LIMIT | ENABLE_MOTOR   (mouse: - target.y target.x -)
//   pulley: j_s0.x = impulse; tmp rows r_a r_b | u_a u_b | mass.   mouse: j_s0.xy = impulse; tmp rows r_a r_b | mass 2x2 | gamma C.x C.y
// and the per-step solver data (the reference's "solver temp" members) in JT_Q float4 rows of j_tmp:
//   0: rA.xy rB.xy
//   1: revolute K.ex.x K.ey.x K.ey.y axial_mass      | distance u.x u.y mass soft_mass        | weld mass.ex.xyz mass.ey.x
//   2: revolute angle - - -                          | distance gamma bias current_length -   | weld mass.ey.yz mass.ez.xy
//   3: mA iA mB iB
//   4: weld mass.ez.z gamma bias -
// prismatic: 0: axis.xy perp.xy   1: s1 s2 a1 a2   2: K11 K12 K22 axial_mass   3: mA iA mB iB   4: translation - - -
// wheel (private joints/b2_wheel_joint.rs:19-170 / :172-282 / :284-380; j_s0 = impulse, spring_impulse, motor_impulse,
//        lower_impulse; param 0 = stiffness, 1 / 2 = translation limits, 5 / 6 = local x axis (not normalised), 7 = damping):
//        0: ax.xy ay.xy   1: sAx sBx sAy sBy   2: mass axial_mass spring_mass motor_mass   3: mA iA mB iB   4: bias gamma translation -
// friction / motor (private joints/b2_friction_joint.rs:8-72 / :74-128, b2_motor_joint.rs:8-96 / :98-160; no position rows;
//        j_s0 = linear impulse x y, angular impulse; param 0 = max force, 1 = max torque, motor: 2 = angular offset,
//        3 = correction factor, local_anchor_a = linear offset):
//        0: rA.xy rB.xy   1: linear mass ex.x ex.y ey.x ey.y   2: angular mass, linear error x y, angular error   3: mA iA mB iB
// Joint visits are ordered work: they run in the island's joint order inside every form of the Gauss-Seidel stages
// (VelocityK / PositionK generic; velocity_sl_kernel / position_sl_kernel for batches, through an accessor over their
// shared-memory rows; LwVelocity7K / LwPosition6K in the large-world modes).
#pragma once
#include "b2g_common.h"

namespace b2g {

#define B2G_ANGULAR_SLOP (2.0f / 180.0f * B2G_PI)            // src/b2_common.rs:43
#define B2G_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2G_PI)  // src/b2_common.rs:64

B2G_HD int jt_at(const Batch& B, const WIdx& x, int j, int q) { return x.at(B.NJ * JT_Q, j * JT_Q + q); }

// B2Mat22::solve (src/b2_math.rs:276-293) for the symmetric K = [ex.x ey.x; ey.x ey.y]
B2G_HD V2 mat22_solve(float exx, float eyx, float exy, float eyy, V2 b) {
  const float a11 = exx, a12 = eyx, a21 = exy, a22 = eyy;
  float det = a11 * a22 - a12 * a21;
  if (det != 0.0f) det = 1.0f / det;
  return v2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}

// B2Mat33 (src/b2_math.rs:296-352, src/private/common/b2_math.rs:5-72) on a plain array: ex.xyz, ey.xyz, ez.xyz at [0..8].
B2G_HD float vec3_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
B2G_HD void vec3_cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
B2G_HD void mat33_get_inverse22(const float* k, float* m) {
  const float a = k[0], b = k[3], c = k[1], d = k[4];
  float det = a * d - b * c;
  if (det != 0.0f) det = 1.0f / det;
  m[0] = det * d; m[3] = -det * b; m[2] = 0.0f;
  m[1] = -det * c; m[4] = det * a; m[5] = 0.0f;
  m[6] = 0.0f; m[7] = 0.0f; m[8] = 0.0f;
}
B2G_HD void mat33_get_sym_inverse33(const float* k, float* m) {
  float cr[3];
  vec3_cross(k + 3, k + 6, cr);
  float det = vec3_dot(k, cr);
  if (det != 0.0f) det = 1.0f / det;
  const float a11 = k[0], a12 = k[3], a13 = k[6], a22 = k[4], a23 = k[7], a33 = k[8];
  m[0] = det * (a22 * a33 - a23 * a23);
  m[1] = det * (a13 * a23 - a12 * a33);
  m[2] = det * (a12 * a23 - a13 * a22);
  m[3] = m[1];
  m[4] = det * (a11 * a33 - a13 * a13);
  m[5] = det * (a13 * a12 - a11 * a23);
  m[6] = m[2];
  m[7] = m[5];
  m[8] = det * (a11 * a22 - a12 * a12);
}
B2G_HD void mat33_solve33(const float* k, const float* b, float* x) {
  float cr[3];
  vec3_cross(k + 3, k + 6, cr);
  float det = vec3_dot(k, cr);
  if (det != 0.0f) det = 1.0f / det;
  x[0] = det * vec3_dot(b, cr);
  vec3_cross(b, k + 6, cr);
  x[1] = det * vec3_dot(k, cr);
  vec3_cross(k + 3, b, cr);
  x[2] = det * vec3_dot(k, cr);
}
// K of the weld joint (b2_weld_joint.rs:66-75 and :238-247: the same expression in both places)
B2G_HD void weld_k(float* k, V2 r_a, V2 r_b, float m_a, float i_a, float m_b, float i_b) {
  k[0] = m_a + m_b + r_a.y * r_a.y * i_a + r_b.y * r_b.y * i_b;
  k[3] = -r_a.y * r_a.x * i_a - r_b.y * r_b.x * i_b;
  k[6] = -r_a.y * i_a - r_b.y * i_b;
  k[1] = k[3];
  k[4] = m_a + m_b + r_a.x * r_a.x * i_a + r_b.x * r_b.x * i_b;
  k[7] = r_a.x * i_a + r_b.x * i_b;
  k[2] = k[6];
  k[5] = k[7];
  k[8] = i_a + i_b;
}

// Does a joint of `self_` prevent collision with `other`? (B2body::should_collide, b2_body.rs(private):400-413)
B2G_HD bool joints_prevent_collision(const Batch& B, int self_, int other) {
  if (B.NJ == 0) return false;
  for (int e = B.jadj_off[self_]; e < B.jadj_off[self_ + 1]; ++e) {
    const int je = B.jadj[e];
    const b2gpu_joint_rec& jr = B.joints[je >> 1];
    const int jn_other = (je & 1) ? jr.body_a : jr.body_b;
    if (jn_other == other && !(jr.flags & B2GPU_JOINT_COLLIDE_CONNECTED)) return true;
  }
  return false;
}

// Where a joint visit reads and writes the running body state.  BodyStateGlobal: the per-world arrays in HBM (generic
// stages, host simulator).  The shared-memory Gauss-Seidel kernels of b2g_solver_smem.cuh pass accessors over their
// [body][lane] rows instead, so joint rows run inside velocity_sl_kernel / position_sl_kernel.
struct BodyStateGlobal {
  const Batch& B;
  const WIdx& x;
  B2G_HD float4 vel(int b) const { return B.b_vel[x.at(B.NB, b)]; }
  B2G_HD void set_vel(int b, float4 v) const { B.b_vel[x.at(B.NB, b)] = v; }
  B2G_HD float4 pos(int b) const { return B.b_pos[x.at(B.NB, b)]; }   // c.x c.y a sleep_time
  B2G_HD void set_pos(int b, float4 p) const { B.b_pos[x.at(B.NB, b)] = p; }
  B2G_HD Rot rot(int b) const { const float4 r = B.b_rot[x.at(B.NB, b)]; Rot q; q.s = r.x; q.c = r.y; return q; }
  B2G_HD void set_rot(int b, Rot q) const { float4 r = B.b_rot[x.at(B.NB, b)]; r.x = q.s; r.y = q.c; B.b_rot[x.at(B.NB, b)] = r; }
};

// ---------------------------------------------------------------------------------------------------------------------
// Gear joint (private joints/b2_gear_joint.rs:142-245 / :247-290 / :292-400): one row over FOUR bodies — A, B (the joint's
// own) and C, D (body A of the two coupled revolute / prismatic joints; ids in impulse[5], [6] of the record).  Every body is
// read first and written back in the order A, B, C, D as the reference does, so that when two of them are one body the later
// store wins.  j_s0.x = impulse (not rescaled by dt_ratio); scratch rows 0: jv_ac jv_bd, 1: jw_a jw_b jw_c jw_d, 2: mass,
// 3: mA iA mB iB, 4: mC iC mD iD.
struct GearBodies {
  int b[4];
  float m[4], i[4];
  V2 lc[4];
};
struct GearRows {
  V2 jv_ac, jv_bd;
  float jw_a, jw_b, jw_c, jw_d, mass, coordinate_a, coordinate_b;
};
B2G_HD GearBodies gear_bodies(const Batch& B, const WIdx& x, const b2gpu_joint_rec& jr) {
  GearBodies g;
  g.b[0] = jr.body_a; g.b[1] = jr.body_b; g.b[2] = f2i(jr.impulse[5]); g.b[3] = f2i(jr.impulse[6]);
  for (int k = 0; k < 4; ++k) {
    const float4 ms = B.b_mass[x.at(B.NB, g.b[k])];
    g.m[k] = ms.x; g.i[k] = ms.y; g.lc[k] = v2(ms.z, ms.w);
  }
  return g;
}
B2G_HD GearRows gear_rows(const b2gpu_joint_rec& jr, const GearBodies& gb, const V2* c, const float* a, const Rot* q) {
  GearRows g;
  const float ratio = jr.impulse[4];
  const V2 anchor_a = v2(jr.local_anchor_a[0], jr.local_anchor_a[1]), anchor_b = v2(jr.local_anchor_b[0], jr.local_anchor_b[1]);
  const V2 anchor_c = v2(jr.param[0], jr.param[1]), anchor_d = v2(jr.param[2], jr.param[3]);
  const V2 axis_c = v2(jr.param[4], jr.param[5]), axis_d = v2(jr.param[6], jr.param[7]);
  g.mass = 0.0f;
  if (!(jr.flags & B2GPU_JOINT_GEAR_PRISMATIC_1)) {
    g.jv_ac = v2(0.0f, 0.0f);
    g.jw_a = 1.0f; g.jw_c = 1.0f;
    g.mass = g.mass + (gb.i[0] + gb.i[2]);
    g.coordinate_a = a[0] - a[2] - jr.impulse[1];
  } else {
    const V2 u = rot_mul(q[2], axis_c);
    const V2 r_c = rot_mul(q[2], anchor_c - gb.lc[2]);
    const V2 r_a = rot_mul(q[0], anchor_a - gb.lc[0]);
    g.jv_ac = u;
    g.jw_c = cross(r_c, u);
    g.jw_a = cross(r_a, u);
    g.mass = g.mass + (gb.m[2] + gb.m[0] + gb.i[2] * g.jw_c * g.jw_c + gb.i[0] * g.jw_a * g.jw_a);
    const V2 p_c = anchor_c - gb.lc[2];
    const V2 p_a = rot_mul_t(q[2], r_a + (c[0] - c[2]));
    g.coordinate_a = dot(p_a - p_c, axis_c);
  }
  if (!(jr.flags & B2GPU_JOINT_GEAR_PRISMATIC_2)) {
    g.jv_bd = v2(0.0f, 0.0f);
    g.jw_b = ratio; g.jw_d = ratio;
    g.mass = g.mass + ratio * ratio * (gb.i[1] + gb.i[3]);
    g.coordinate_b = a[1] - a[3] - jr.impulse[2];
  } else {
    const V2 u = rot_mul(q[3], axis_d);
    const V2 r_d = rot_mul(q[3], anchor_d - gb.lc[3]);
    const V2 r_b = rot_mul(q[1], anchor_b - gb.lc[1]);
    g.jv_bd = ratio * u;
    g.jw_d = ratio * cross(r_d, u);
    g.jw_b = ratio * cross(r_b, u);
    g.mass = g.mass + (ratio * ratio * (gb.m[3] + gb.m[1]) + gb.i[3] * g.jw_d * g.jw_d + gb.i[1] * g.jw_b * g.jw_b);
    const V2 p_d = anchor_d - gb.lc[3];
    const V2 p_b = rot_mul_t(q[3], r_b + (c[1] - c[3]));
    g.coordinate_b = dot(p_b - p_d, axis_d);
  }
  return g;
}
template <class S>
B2G_HD void gear_init_velocity(const Batch& B, const WIdx& x, const S& st, int j, bool warm_starting) {
  const b2gpu_joint_rec& jr = B.joints[j];
  const GearBodies gb = gear_bodies(B, x, jr);
  V2 c[4], v[4];
  float a[4], w[4];
  Rot q[4];
  for (int k = 0; k < 4; ++k) {
    const float4 p = st.pos(gb.b[k]), vl = st.vel(gb.b[k]);
    c[k] = v2(p.x, p.y); a[k] = p.z;
    v[k] = v2(vl.x, vl.y); w[k] = vl.z;
    q[k] = st.rot(gb.b[k]);
  }
  const GearRows g = gear_rows(jr, gb, c, a, q);
  const float mass = g.mass > 0.0f ? 1.0f / g.mass : 0.0f;
  const int ji = x.at(B.NJ, j);
  float4 s0 = B.j_s0[ji];
  if (warm_starting) {
    const float imp = s0.x;
    v[0] = v[0] + (gb.m[0] * imp) * g.jv_ac;
    w[0] += gb.i[0] * imp * g.jw_a;
    v[1] = v[1] + (gb.m[1] * imp) * g.jv_bd;
    w[1] += gb.i[1] * imp * g.jw_b;
    v[2] = v[2] - (gb.m[2] * imp) * g.jv_ac;
    w[2] -= gb.i[2] * imp * g.jw_c;
    v[3] = v[3] - (gb.m[3] * imp) * g.jv_bd;
    w[3] -= gb.i[3] * imp * g.jw_d;
  } else {
    s0.x = 0.0f;
  }
  B.j_s0[ji] = s0;
  B.j_tmp[jt_at(B, x, j, 0)] = make_float4(g.jv_ac.x, g.jv_ac.y, g.jv_bd.x, g.jv_bd.y);
  B.j_tmp[jt_at(B, x, j, 1)] = make_float4(g.jw_a, g.jw_b, g.jw_c, g.jw_d);
  B.j_tmp[jt_at(B, x, j, 2)] = make_float4(mass, 0.0f, 0.0f, 0.0f);
  B.j_tmp[jt_at(B, x, j, 3)] = make_float4(gb.m[0], gb.i[0], gb.m[1], gb.i[1]);
  B.j_tmp[jt_at(B, x, j, 4)] = make_float4(gb.m[2], gb.i[2], gb.m[3], gb.i[3]);
  for (int k = 0; k < 4; ++k)
    if (gb.m[k] != 0.0f || gb.i[k] != 0.0f) st.set_vel(gb.b[k], make_float4(v[k].x, v[k].y, w[k], 0.0f));
}
template <class S>
B2G_HD void gear_solve_velocity(const Batch& B, const WIdx& x, const S& st, int j) {
  const b2gpu_joint_rec& jr = B.joints[j];
  const int b[4] = {jr.body_a, jr.body_b, f2i(jr.impulse[5]), f2i(jr.impulse[6])};
  const float4 t0 = B.j_tmp[jt_at(B, x, j, 0)], t1 = B.j_tmp[jt_at(B, x, j, 1)], t2 = B.j_tmp[jt_at(B, x, j, 2)];
  const float4 t3 = B.j_tmp[jt_at(B, x, j, 3)], t4 = B.j_tmp[jt_at(B, x, j, 4)];
  const float m[4] = {t3.x, t3.z, t4.x, t4.z}, i[4] = {t3.y, t3.w, t4.y, t4.w};
  const V2 jv_ac = v2(t0.x, t0.y), jv_bd = v2(t0.z, t0.w);
  V2 v[4];
  float w[4];
  for (int k = 0; k < 4; ++k) {
    const float4 vl = st.vel(b[k]);
    v[k] = v2(vl.x, vl.y); w[k] = vl.z;
  }
  float cdot = dot(jv_ac, v[0] - v[2]) + dot(jv_bd, v[1] - v[3]);
  cdot += (t1.x * w[0] - t1.z * w[2]) + (t1.y * w[1] - t1.w * w[3]);
  const float impulse = -t2.x * cdot;
  const int ji = x.at(B.NJ, j);
  B.j_s0[ji].x += impulse;
  v[0] = v[0] + (m[0] * impulse) * jv_ac;
  w[0] += i[0] * impulse * t1.x;
  v[1] = v[1] + (m[1] * impulse) * jv_bd;
  w[1] += i[1] * impulse * t1.y;
  v[2] = v[2] - (m[2] * impulse) * jv_ac;
  w[2] -= i[2] * impulse * t1.z;
  v[3] = v[3] - (m[3] * impulse) * jv_bd;
  w[3] -= i[3] * impulse * t1.w;
  for (int k = 0; k < 4; ++k)
    if (m[k] != 0.0f || i[k] != 0.0f) st.set_vel(b[k], make_float4(v[k].x, v[k].y, w[k], 0.0f));
}
template <class S>
B2G_HD bool gear_solve_position(const Batch& B, const WIdx& x, const S& st, int j) {
  const b2gpu_joint_rec& jr = B.joints[j];
  const GearBodies gb = gear_bodies(B, x, jr);
  V2 c[4];
  float a[4];
  Rot q[4];
  for (int k = 0; k < 4; ++k) {
    const float4 p = st.pos(gb.b[k]);
    c[k] = v2(p.x, p.y); a[k] = p.z;
    q[k] = st.rot(gb.b[k]);
  }
  const GearRows g = gear_rows(jr, gb, c, a, q);
  const float cc = (g.coordinate_a + jr.impulse[4] * g.coordinate_b) - jr.impulse[3];
  float impulse = 0.0f;
  if (g.mass > 0.0f) impulse = -cc / g.mass;
  c[0] = c[0] + (gb.m[0] * impulse) * g.jv_ac;
  a[0] += gb.i[0] * impulse * g.jw_a;
  c[1] = c[1] + (gb.m[1] * impulse) * g.jv_bd;
  a[1] += gb.i[1] * impulse * g.jw_b;
  c[2] = c[2] - (gb.m[2] * impulse) * g.jv_ac;
  a[2] -= gb.i[2] * impulse * g.jw_c;
  c[3] = c[3] - (gb.m[3] * impulse) * g.jv_bd;
  a[3] -= gb.i[3] * impulse * g.jw_d;
  for (int k = 0; k < 4; ++k) {
    if (gb.m[k] == 0.0f && gb.i[k] == 0.0f) continue;  // immovable bodies are shared between islands: never written
    float4 cur = st.pos(gb.b[k]);                      // what an earlier store of this loop may have left (aliased bodies)
    if (f2u(a[k]) != f2u(cur.z)) st.set_rot(gb.b[k], rot_from_angle(a[k]));
    cur.x = c[k].x; cur.y = c[k].y; cur.z = a[k];
    st.set_pos(gb.b[k], cur);
  }
  return true;  // linear_error stays 0 (b2_gear_joint.rs:311, :399)
}

// init_velocity_constraints of joint j: solver data into j_tmp, impulses scaled (or zeroed) in j_s0 / j_s1, the warm-start
// impulse applied to the two bodies' velocities.
template <class S>
B2G_HD void joint_init_velocity(const Batch& B, const WIdx& x, const S& st, int j, bool warm_starting, float dt_ratio, float h) {
  const b2gpu_joint_rec& jr = B.joints[j];
  if (jr.type == B2GPU_JOINT_GEAR) { gear_init_velocity(B, x, st, j, warm_starting); return; }
  const int bai = x.at(B.NB, jr.body_a), bbi = x.at(B.NB, jr.body_b);
  const float4 msa = B.b_mass[bai], msb = B.b_mass[bbi];
  const float m_a = msa.x, i_a = msa.y, m_b = msb.x, i_b = msb.y;
  const float4 pa = st.pos(jr.body_a), pb = st.pos(jr.body_b);
  const Rot rqa = st.rot(jr.body_a), rqb = st.rot(jr.body_b);  // B2Rot::new(a) of the island body's angle (IntegrateK)
  const float4 ra4 = make_float4(rqa.s, rqa.c, 0.0f, 0.0f), rb4 = make_float4(rqb.s, rqb.c, 0.0f, 0.0f);
  float4 va = st.vel(jr.body_a), vb = st.vel(jr.body_b);
  V2 v_a = v2(va.x, va.y), v_b = v2(vb.x, vb.y);
  float w_a = va.z, w_b = vb.z;
  const V2 c_a = v2(pa.x, pa.y), c_b = v2(pb.x, pb.y);
  const float a_a = pa.z, a_b = pb.z;
  Rot q_a, q_b;
  q_a.s = ra4.x; q_a.c = ra4.y;
  q_b.s = rb4.x; q_b.c = rb4.y;
  const V2 r_a = rot_mul(q_a, v2(jr.local_anchor_a[0], jr.local_anchor_a[1]) - v2(msa.z, msa.w));
  const V2 r_b = rot_mul(q_b, v2(jr.local_anchor_b[0], jr.local_anchor_b[1]) - v2(msb.z, msb.w));
  const int ji = x.at(B.NJ, j);
  float4 s0 = B.j_s0[ji], s1 = B.j_s1[ji];
  const int jflags = f2i(s1.w);
  float4 t1, t2 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), t4 = t2;
  float4 t0 = make_float4(r_a.x, r_a.y, r_b.x, r_b.y);
  if (jr.type == B2GPU_JOINT_FRICTION || jr.type == B2GPU_JOINT_MOTOR) {
    V2 fr_a = r_a, fr_b = r_b;
    if (jr.type == B2GPU_JOINT_MOTOR) fr_b = rot_mul(q_b, -v2(msb.z, msb.w));  // from body B's origin: -local_center_b, not 0 - it
    const float kxx = m_a + m_b + i_a * fr_a.y * fr_a.y + i_b * fr_b.y * fr_b.y;
    const float kxy = -i_a * fr_a.x * fr_a.y - i_b * fr_b.x * fr_b.y;
    const float kyy = m_a + m_b + i_a * fr_a.x * fr_a.x + i_b * fr_b.x * fr_b.x;
    // B2Mat22::get_inverse (src/b2_math.rs:261-274) of [kxx kxy; kxy kyy]
    float det = kxx * kyy - kxy * kxy;
    if (det != 0.0f) det = 1.0f / det;
    t1 = make_float4(det * kyy, -det * kxy, -det * kxy, det * kxx);
    float angular_mass = i_a + i_b;
    if (angular_mass > 0.0f) angular_mass = 1.0f / angular_mass;
    V2 linear_error = v2(0.0f, 0.0f);
    float angular_error = 0.0f;
    if (jr.type == B2GPU_JOINT_MOTOR) {
      linear_error = c_b + fr_b - c_a - fr_a;
      angular_error = a_b - a_a - jr.param[2];
    }
    if (warm_starting) {
      s0.x *= dt_ratio; s0.y *= dt_ratio;
      s0.z *= dt_ratio;
      const V2 p = v2(s0.x, s0.y);
      v_a = v_a - m_a * p;
      w_a -= i_a * (cross(fr_a, p) + s0.z);
      v_b = v_b + m_b * p;
      w_b += i_b * (cross(fr_b, p) + s0.z);
    } else {
      s0.x = 0.0f; s0.y = 0.0f; s0.z = 0.0f;
    }
    t0 = make_float4(fr_a.x, fr_a.y, fr_b.x, fr_b.y);
    t2 = make_float4(angular_mass, linear_error.x, linear_error.y, angular_error);
  } else if (jr.type == B2GPU_JOINT_PULLEY) {  // private joints/b2_pulley_joint.rs:46-137
    const float ratio = jr.param[6];
    V2 u_a = c_a + r_a - v2(jr.param[0], jr.param[1]);
    V2 u_b = c_b + r_b - v2(jr.param[2], jr.param[3]);
    const float length_a = length(u_a), length_b = length(u_b);
    if (length_a > 10.0f * B2G_LINEAR_SLOP) u_a = (1.0f / length_a) * u_a; else u_a = v2(0.0f, 0.0f);
    if (length_b > 10.0f * B2G_LINEAR_SLOP) u_b = (1.0f / length_b) * u_b; else u_b = v2(0.0f, 0.0f);
    const float ru_a = cross(r_a, u_a), ru_b = cross(r_b, u_b);
    const float pm_a = m_a + i_a * ru_a * ru_a;
    const float pm_b = m_b + i_b * ru_b * ru_b;
    float mass = pm_a + ratio * ratio * pm_b;
    if (mass > 0.0f) mass = 1.0f / mass;
    if (warm_starting) {
      s0.x *= dt_ratio;
      const V2 pa_ = (-s0.x) * u_a;
      const V2 pb_ = (-ratio * s0.x) * u_b;
      v_a = v_a + m_a * pa_;
      w_a += i_a * cross(r_a, pa_);
      v_b = v_b + m_b * pb_;
      w_b += i_b * cross(r_b, pb_);
    } else {
      s0.x = 0.0f;
    }
    t1 = make_float4(u_a.x, u_a.y, u_b.x, u_b.y);
    t2 = make_float4(mass, 0.0f, 0.0f, 0.0f);
  } else if (jr.type == B2GPU_JOINT_MOUSE) {  // private joints/b2_mouse_joint.rs:7-67: body A is neither read nor changed
    const float d = jr.param[2], k = jr.param[1];
    float gamma = h * (d + h * k);
    if (gamma != 0.0f) gamma = 1.0f / gamma;
    const float beta = h * k * gamma;
    const float kxx = m_b + i_b * r_b.y * r_b.y + gamma;
    const float kxy = -i_b * r_b.x * r_b.y;
    const float kyy = m_b + i_b * r_b.x * r_b.x + gamma;
    float det = kxx * kyy - kxy * kxy;  // B2Mat22::get_inverse (src/b2_math.rs:261-274)
    if (det != 0.0f) det = 1.0f / det;
    t1 = make_float4(det * kyy, -det * kxy, -det * kxy, det * kxx);
    V2 c = c_b + r_b - v2(s1.z, s1.y);  // the target lives in j_s1 (per world)
    c = beta * c;
    w_b *= 0.98f;
    if (warm_starting) {
      s0.x *= dt_ratio; s0.y *= dt_ratio;
      const V2 p = v2(s0.x, s0.y);
      v_b = v_b + m_b * p;
      w_b += i_b * cross(r_b, p);
    } else {
      s0.x = 0.0f; s0.y = 0.0f;
    }
    t2 = make_float4(gamma, c.x, c.y, 0.0f);
  } else if (jr.type == B2GPU_JOINT_WHEEL) {
    const V2 d = c_b + r_b - c_a - r_a;
    const V2 lx = v2(jr.param[5], jr.param[6]), ly = cross_sv(1.0f, lx);
    const V2 ay = rot_mul(q_a, ly);
    const float s_ay = cross(d + r_a, ay), s_by = cross(r_b, ay);
    float mass = m_a + m_b + i_a * s_ay * s_ay + i_b * s_by * s_by;
    if (mass > 0.0f) mass = 1.0f / mass;
    const V2 ax = rot_mul(q_a, lx);
    const float s_ax = cross(d + r_a, ax), s_bx = cross(r_b, ax);
    const float inv_mass = m_a + m_b + i_a * s_ax * s_ax + i_b * s_bx * s_bx;
    const float axial_mass = inv_mass > 0.0f ? 1.0f / inv_mass : 0.0f;
    float spring_mass = 0.0f, bias = 0.0f, gamma = 0.0f;
    const float stiffness = jr.param[0], damping = jr.param[7];
    if (stiffness > 0.0f && inv_mass > 0.0f) {
      spring_mass = 1.0f / inv_mass;
      const float c = dot(d, ax);
      gamma = h * (damping + h * stiffness);
      if (gamma > 0.0f) gamma = 1.0f / gamma;
      bias = c * h * stiffness * gamma;
      spring_mass = inv_mass + gamma;
      if (spring_mass > 0.0f) spring_mass = 1.0f / spring_mass;
    } else {
      s0.y = 0.0f;
    }
    float translation = 0.0f;
    if (jflags & B2GPU_JOINT_ENABLE_LIMIT) translation = dot(ax, d);
    else { s0.w = 0.0f; s1.x = 0.0f; }
    float motor_mass;
    if (jflags & B2GPU_JOINT_ENABLE_MOTOR) {
      motor_mass = i_a + i_b;
      if (motor_mass > 0.0f) motor_mass = 1.0f / motor_mass;
    } else {
      motor_mass = 0.0f;
      s0.z = 0.0f;
    }
    if (warm_starting) {
      s0.x *= dt_ratio;  // the limit impulses are not scaled (b2_wheel_joint.rs:143-146)
      s0.y *= dt_ratio;
      s0.z *= dt_ratio;
      const float axial_impulse = s0.y + s0.w - s1.x;
      const V2 p = s0.x * ay + axial_impulse * ax;
      const float la = s0.x * s_ay + axial_impulse * s_ax + s0.z;
      const float lb = s0.x * s_by + axial_impulse * s_bx + s0.z;
      v_a = v_a - m_a * p;
      w_a -= i_a * la;
      v_b = v_b + m_b * p;
      w_b += i_b * lb;
    } else {
      s0.x = 0.0f; s0.y = 0.0f; s0.z = 0.0f; s0.w = 0.0f; s1.x = 0.0f;
    }
    t0 = make_float4(ax.x, ax.y, ay.x, ay.y);
    t1 = make_float4(s_ax, s_bx, s_ay, s_by);
    t2 = make_float4(mass, axial_mass, spring_mass, motor_mass);
    t4 = make_float4(bias, gamma, translation, 0.0f);
  } else if (jr.type == B2GPU_JOINT_PRISMATIC) {
    const V2 d = (c_b - c_a) + r_b - r_a;
    const V2 lx = v2(jr.param[5], jr.param[6]), ly = cross_sv(1.0f, lx);  // m_local_xaxis_a, m_local_yaxis_a
    const V2 axis = rot_mul(q_a, lx);
    const float a1 = cross(d + r_a, axis), a2 = cross(r_b, axis);
    float axial_mass = m_a + m_b + i_a * a1 * a1 + i_b * a2 * a2;
    if (axial_mass > 0.0f) axial_mass = 1.0f / axial_mass;
    const V2 perp = rot_mul(q_a, ly);
    const float ps1 = cross(d + r_a, perp), ps2 = cross(r_b, perp);
    const float k11 = m_a + m_b + i_a * ps1 * ps1 + i_b * ps2 * ps2;
    const float k12 = i_a * ps1 + i_b * ps2;
    float k22 = i_a + i_b;
    if (k22 == 0.0f) k22 = 1.0f;  // bodies with fixed rotation
    float translation = 0.0f;
    if (jflags & B2GPU_JOINT_ENABLE_LIMIT) translation = dot(axis, d);
    else { s0.w = 0.0f; s1.x = 0.0f; }
    if (!(jflags & B2GPU_JOINT_ENABLE_MOTOR)) s0.z = 0.0f;
    if (warm_starting) {
      s0.x *= dt_ratio; s0.y *= dt_ratio;
      s0.z *= dt_ratio;
      s0.w *= dt_ratio;
      s1.x *= dt_ratio;
      const float axial_impulse = s0.z + s0.w - s1.x;
      const V2 p = s0.x * perp + axial_impulse * axis;
      const float la = s0.x * ps1 + s0.y + axial_impulse * a1;
      const float lb = s0.x * ps2 + s0.y + axial_impulse * a2;
      v_a = v_a - m_a * p;
      w_a -= i_a * la;
      v_b = v_b + m_b * p;
      w_b += i_b * lb;
    } else {
      s0.x = 0.0f; s0.y = 0.0f; s0.z = 0.0f; s0.w = 0.0f; s1.x = 0.0f;
    }
    t0 = make_float4(axis.x, axis.y, perp.x, perp.y);
    t1 = make_float4(ps1, ps2, a1, a2);
    t2 = make_float4(k11, k12, k22, axial_mass);
    t4 = make_float4(translation, 0.0f, 0.0f, 0.0f);
  } else if (jr.type == B2GPU_JOINT_WELD) {
    float k[9], m[9];
    weld_k(k, r_a, r_b, m_a, i_a, m_b, i_b);
    const float stiffness = jr.param[3], damping = jr.param[4];
    float gamma, bias;
    if (stiffness > 0.0f) {
      mat33_get_inverse22(k, m);
      float inv_m = i_a + i_b;
      const float c = a_b - a_a - jr.param[0];
      gamma = h * (damping + h * stiffness);
      gamma = gamma != 0.0f ? 1.0f / gamma : 0.0f;
      bias = c * h * stiffness * gamma;
      inv_m += gamma;
      m[8] = inv_m != 0.0f ? 1.0f / inv_m : 0.0f;
    } else if (k[8] == 0.0f) {
      mat33_get_inverse22(k, m);
      gamma = 0.0f;
      bias = 0.0f;
    } else {
      mat33_get_sym_inverse33(k, m);
      gamma = 0.0f;
      bias = 0.0f;
    }
    if (warm_starting) {
      s0.x *= dt_ratio; s0.y *= dt_ratio; s0.z *= dt_ratio;
      const V2 p = v2(s0.x, s0.y);
      v_a = v_a - m_a * p;
      w_a -= i_a * (cross(r_a, p) + s0.z);
      v_b = v_b + m_b * p;
      w_b += i_b * (cross(r_b, p) + s0.z);
    } else {
      s0.x = 0.0f; s0.y = 0.0f; s0.z = 0.0f;
    }
    t1 = make_float4(m[0], m[1], m[2], m[3]);
    t2 = make_float4(m[4], m[5], m[6], m[7]);
    t4 = make_float4(m[8], gamma, bias, 0.0f);
  } else if (jr.type == B2GPU_JOINT_REVOLUTE) {
    const float kxx = m_a + m_b + r_a.y * r_a.y * i_a + r_b.y * r_b.y * i_b;
    const float kyx = -r_a.y * r_a.x * i_a - r_b.y * r_b.x * i_b;
    const float kyy = m_a + m_b + r_a.x * r_a.x * i_a + r_b.x * r_b.x * i_b;
    float axial_mass = i_a + i_b;
    bool fixed_rotation;
    if (axial_mass > 0.0f) { axial_mass = 1.0f / axial_mass; fixed_rotation = false; }
    else fixed_rotation = true;
    t2.x = a_b - a_a - jr.param[0];
    if (!(jflags & B2GPU_JOINT_ENABLE_LIMIT) || fixed_rotation) { s0.w = 0.0f; s1.x = 0.0f; }
    if (!(jflags & B2GPU_JOINT_ENABLE_MOTOR) || fixed_rotation) s0.z = 0.0f;
    if (warm_starting) {
      s0.x *= dt_ratio; s0.y *= dt_ratio;
      s0.z *= dt_ratio;
      s0.w *= dt_ratio;
      s1.x *= dt_ratio;
      const float axial_impulse = s0.z + s0.w - s1.x;
      const V2 p = v2(s0.x, s0.y);
      v_a = v_a - m_a * p;
      w_a -= i_a * (cross(r_a, p) + axial_impulse);
      v_b = v_b + m_b * p;
      w_b += i_b * (cross(r_b, p) + axial_impulse);
    } else {
      s0.x = 0.0f; s0.y = 0.0f; s0.z = 0.0f; s0.w = 0.0f; s1.x = 0.0f;
    }
    t1 = make_float4(kxx, kyx, kyy, axial_mass);
  } else {  // distance
    V2 u = c_b + r_b - c_a - r_a;
    const float current_length = length(u);
    if (current_length > B2G_LINEAR_SLOP) {
      u = (1.0f / current_length) * u;
    } else {
      u = v2(0.0f, 0.0f);
      s0.x = 0.0f; s0.w = 0.0f; s1.x = 0.0f;
    }
    const float cr_au = cross(r_a, u), cr_bu = cross(r_b, u);
    float inv_mass = m_a + i_a * cr_au * cr_au + m_b + i_b * cr_bu * cr_bu;
    const float mass = inv_mass != 0.0f ? 1.0f / inv_mass : 0.0f;
    float soft_mass, gamma, bias;
    const float length_ = jr.param[0], min_length = jr.param[1], max_length = jr.param[2], stiffness = jr.param[3], damping = jr.param[4];
    if (stiffness > 0.0f && min_length < max_length) {
      const float c = current_length - length_;
      gamma = h * (damping + h * stiffness);
      gamma = gamma != 0.0f ? 1.0f / gamma : 0.0f;
      bias = c * h * stiffness * gamma;
      inv_mass += gamma;
      soft_mass = inv_mass != 0.0f ? 1.0f / inv_mass : 0.0f;
    } else {
      gamma = 0.0f;
      bias = 0.0f;
      soft_mass = mass;
    }
    if (warm_starting) {
      s0.x *= dt_ratio;
      s0.w *= dt_ratio;
      s1.x *= dt_ratio;
      const V2 p = (s0.x + s0.w - s1.x) * u;
      v_a = v_a - m_a * p;
      w_a -= i_a * cross(r_a, p);
      v_b = v_b + m_b * p;
      w_b += i_b * cross(r_b, p);
    } else {
      s0.x = 0.0f;
    }
    t1 = make_float4(u.x, u.y, mass, soft_mass);
    t2 = make_float4(gamma, bias, current_length, 0.0f);
  }
  B.j_s0[ji] = s0;
  B.j_s1[ji] = s1;
  B.j_tmp[jt_at(B, x, j, 0)] = t0;
  B.j_tmp[jt_at(B, x, j, 1)] = t1;
  B.j_tmp[jt_at(B, x, j, 2)] = t2;
  B.j_tmp[jt_at(B, x, j, 3)] = make_float4(m_a, i_a, m_b, i_b);
  if (jr.type == B2GPU_JOINT_WELD || jr.type == B2GPU_JOINT_PRISMATIC || jr.type == B2GPU_JOINT_WHEEL) B.j_tmp[jt_at(B, x, j, 4)] = t4;
  // an immovable body may sit in several islands: its velocity never changes, leave it alone — but for the mouse joint's
  // damping of w_b, which the reference applies to a kinematic body B too (a kinematic body belongs to one island)
  const bool mouse_kinematic = jr.type == B2GPU_JOINT_MOUSE && body_type(B.b_flags[bbi]) != B2GPU_STATIC_BODY;
  if (m_a != 0.0f || i_a != 0.0f) st.set_vel(jr.body_a, make_float4(v_a.x, v_a.y, w_a, 0.0f));
  if (m_b != 0.0f || i_b != 0.0f || mouse_kinematic) st.set_vel(jr.body_b, make_float4(v_b.x, v_b.y, w_b, 0.0f));
}

template <class S>
B2G_HD void joint_solve_velocity(const Batch& B, const WIdx& x, const S& st, int j, float h, float inv_dt) {
  const b2gpu_joint_rec& jr = B.joints[j];
  if (jr.type == B2GPU_JOINT_GEAR) { gear_solve_velocity(B, x, st, j); return; }
  const float4 va = st.vel(jr.body_a), vb = st.vel(jr.body_b);
  V2 v_a = v2(va.x, va.y), v_b = v2(vb.x, vb.y);
  float w_a = va.z, w_b = vb.z;
  const float4 t0 = B.j_tmp[jt_at(B, x, j, 0)], t1 = B.j_tmp[jt_at(B, x, j, 1)], t2 = B.j_tmp[jt_at(B, x, j, 2)];
  const float4 t3 = B.j_tmp[jt_at(B, x, j, 3)];
  const float m_a = t3.x, i_a = t3.y, m_b = t3.z, i_b = t3.w;
  const V2 r_a = v2(t0.x, t0.y), r_b = v2(t0.z, t0.w);
  const int ji = x.at(B.NJ, j);
  float4 s0 = B.j_s0[ji], s1 = B.j_s1[ji];
  const int jflags = f2i(s1.w);
  if (jr.type == B2GPU_JOINT_FRICTION || jr.type == B2GPU_JOINT_MOTOR) {
    const bool motor = jr.type == B2GPU_JOINT_MOTOR;
    const float angular_mass = t2.x;
    {  // angular
      float cdot = w_b - w_a;
      if (motor) cdot = w_b - w_a + inv_dt * jr.param[3] * t2.w;
      float impulse = -angular_mass * cdot;
      const float old_impulse = s0.z;
      const float max_impulse = h * jr.param[1];
      s0.z = fclamp_sel(s0.z + impulse, -max_impulse, max_impulse);
      impulse = s0.z - old_impulse;
      w_a -= i_a * impulse;
      w_b += i_b * impulse;
    }
    {  // linear
      V2 cdot = v_b + cross_sv(w_b, r_b) - v_a - cross_sv(w_a, r_a);
      if (motor) cdot = cdot + (inv_dt * jr.param[3]) * v2(t2.y, t2.z);
      V2 impulse = -v2(t1.x * cdot.x + t1.z * cdot.y, t1.y * cdot.x + t1.w * cdot.y);  // b2_mul(linear mass, cdot)
      const V2 old_impulse = v2(s0.x, s0.y);
      V2 li = old_impulse + impulse;
      const float max_impulse = h * jr.param[0];
      if (dot(li, li) > max_impulse * max_impulse) {
        normalize(li);
        li = max_impulse * li;
      }
      s0.x = li.x; s0.y = li.y;
      impulse = li - old_impulse;
      v_a = v_a - m_a * impulse;
      w_a -= i_a * cross(r_a, impulse);
      v_b = v_b + m_b * impulse;
      w_b += i_b * cross(r_b, impulse);
    }
  } else if (jr.type == B2GPU_JOINT_PULLEY) {  // private joints/b2_pulley_joint.rs:139-170
    const V2 u_a = v2(t1.x, t1.y), u_b = v2(t1.z, t1.w);
    const float ratio = jr.param[6];
    const V2 vp_a = v_a + cross_sv(w_a, r_a);
    const V2 vp_b = v_b + cross_sv(w_b, r_b);
    const float cdot = -dot(u_a, vp_a) - ratio * dot(u_b, vp_b);
    const float impulse = -t2.x * cdot;
    s0.x += impulse;
    const V2 pa_ = (-impulse) * u_a;
    const V2 pb_ = (-ratio * impulse) * u_b;
    v_a = v_a + m_a * pa_;
    w_a += i_a * cross(r_a, pa_);
    v_b = v_b + m_b * pb_;
    w_b += i_b * cross(r_b, pb_);
  } else if (jr.type == B2GPU_JOINT_MOUSE) {  // private joints/b2_mouse_joint.rs:69-92
    const V2 cdot = v_b + cross_sv(w_b, r_b);
    const V2 old_impulse = v2(s0.x, s0.y);
    const V2 t = -(cdot + v2(t2.y, t2.z) + t2.x * old_impulse);
    V2 impulse = v2(t1.x * t.x + t1.z * t.y, t1.y * t.x + t1.w * t.y);  // b2_mul(mass, t)
    V2 li = old_impulse + impulse;
    const float max_impulse = h * jr.param[0];
    if (dot(li, li) > max_impulse * max_impulse) li = (max_impulse / length(li)) * li;
    s0.x = li.x; s0.y = li.y;
    impulse = li - old_impulse;
    v_b = v_b + m_b * impulse;
    w_b += i_b * cross(r_b, impulse);
  } else if (jr.type == B2GPU_JOINT_WHEEL) {
    const V2 ax = v2(t0.x, t0.y), ay = v2(t0.z, t0.w);
    const float s_ax = t1.x, s_bx = t1.y, s_ay = t1.z, s_by = t1.w;
    const float mass = t2.x, axial_mass = t2.y, spring_mass = t2.z, motor_mass = t2.w;
    const float4 t4 = B.j_tmp[jt_at(B, x, j, 4)];
    const float bias = t4.x, gamma = t4.y, translation = t4.z;
    {  // spring constraint
      const float cdot = dot(ax, v_b - v_a) + s_bx * w_b - s_ax * w_a;
      const float impulse = -spring_mass * (cdot + bias + gamma * s0.y);
      s0.y += impulse;
      const V2 p = impulse * ax;
      const float la = impulse * s_ax, lb = impulse * s_bx;
      v_a = v_a - m_a * p;
      w_a -= i_a * la;
      v_b = v_b + m_b * p;
      w_b += i_b * lb;
    }
    {  // rotational motor constraint (motor_mass is 0 when the motor is off)
      const float cdot = w_b - w_a - s1.y;
      float impulse = -motor_mass * cdot;
      const float old_impulse = s0.z;
      const float max_impulse = h * s1.z;
      s0.z = fclamp_sel(s0.z + impulse, -max_impulse, max_impulse);
      impulse = s0.z - old_impulse;
      w_a -= i_a * impulse;
      w_b += i_b * impulse;
    }
    if (jflags & B2GPU_JOINT_ENABLE_LIMIT) {
      {  // lower limit
        const float c = translation - jr.param[1];
        const float cdot = dot(ax, v_b - v_a) + s_bx * w_b - s_ax * w_a;
        float impulse = -axial_mass * (cdot + fmax_sel(c, 0.0f) * inv_dt);
        const float old_impulse = s0.w;
        s0.w = fmax_sel(s0.w + impulse, 0.0f);
        impulse = s0.w - old_impulse;
        const V2 p = impulse * ax;
        const float la = impulse * s_ax, lb = impulse * s_bx;
        v_a = v_a - m_a * p;
        w_a -= i_a * la;
        v_b = v_b + m_b * p;
        w_b += i_b * lb;
      }
      {  // upper limit
        const float c = jr.param[2] - translation;
        const float cdot = dot(ax, v_a - v_b) + s_ax * w_a - s_bx * w_b;
        float impulse = -axial_mass * (cdot + fmax_sel(c, 0.0f) * inv_dt);
        const float old_impulse = s1.x;
        s1.x = fmax_sel(s1.x + impulse, 0.0f);
        impulse = s1.x - old_impulse;
        const V2 p = impulse * ax;
        const float la = impulse * s_ax, lb = impulse * s_bx;
        v_a = v_a + m_a * p;
        w_a += i_a * la;
        v_b = v_b - m_b * p;
        w_b -= i_b * lb;
      }
    }
    {  // point to line constraint
      const float cdot = dot(ay, v_b - v_a) + s_by * w_b - s_ay * w_a;
      const float impulse = -mass * cdot;
      s0.x += impulse;
      const V2 p = impulse * ay;
      const float la = impulse * s_ay, lb = impulse * s_by;
      v_a = v_a - m_a * p;
      w_a -= i_a * la;
      v_b = v_b + m_b * p;
      w_b += i_b * lb;
    }
  } else if (jr.type == B2GPU_JOINT_PRISMATIC) {
    const V2 axis = v2(t0.x, t0.y), perp = v2(t0.z, t0.w);
    const float s1_ = t1.x, s2_ = t1.y, a1 = t1.z, a2 = t1.w, axial_mass = t2.w;
    const float translation = B.j_tmp[jt_at(B, x, j, 4)].x;
    if (jflags & B2GPU_JOINT_ENABLE_MOTOR) {  // linear motor
      const float cdot = dot(axis, v_b - v_a) + a2 * w_b - a1 * w_a;
      float impulse = axial_mass * (s1.y - cdot);
      const float old_impulse = s0.z;
      const float max_impulse = h * s1.z;
      s0.z = fclamp_sel(s0.z + impulse, -max_impulse, max_impulse);
      impulse = s0.z - old_impulse;
      const V2 p = impulse * axis;
      const float la = impulse * a1, lb = impulse * a2;
      v_a = v_a - m_a * p;
      w_a -= i_a * la;
      v_b = v_b + m_b * p;
      w_b += i_b * lb;
    }
    if (jflags & B2GPU_JOINT_ENABLE_LIMIT) {
      {  // lower limit
        const float c = translation - jr.param[1];
        const float cdot = dot(axis, v_b - v_a) + a2 * w_b - a1 * w_a;
        float impulse = -axial_mass * (cdot + fmax_sel(c, 0.0f) * inv_dt);
        const float old_impulse = s0.w;
        s0.w = fmax_sel(s0.w + impulse, 0.0f);
        impulse = s0.w - old_impulse;
        const V2 p = impulse * axis;
        const float la = impulse * a1, lb = impulse * a2;
        v_a = v_a - m_a * p;
        w_a -= i_a * la;
        v_b = v_b + m_b * p;
        w_b += i_b * lb;
      }
      {  // upper limit: signs flipped to keep c positive when the constraint is satisfied
        const float c = jr.param[2] - translation;
        const float cdot = dot(axis, v_a - v_b) + a1 * w_a - a2 * w_b;
        float impulse = -axial_mass * (cdot + fmax_sel(c, 0.0f) * inv_dt);
        const float old_impulse = s1.x;
        s1.x = fmax_sel(s1.x + impulse, 0.0f);
        impulse = s1.x - old_impulse;
        const V2 p = impulse * axis;
        const float la = impulse * a1, lb = impulse * a2;
        v_a = v_a + m_a * p;
        w_a += i_a * la;
        v_b = v_b - m_b * p;
        w_b -= i_b * lb;
      }
    }
    {  // prismatic constraint in 2D
      const V2 cdot = v2(dot(perp, v_b - v_a) + s2_ * w_b - s1_ * w_a, w_b - w_a);
      const V2 df = mat22_solve(t2.x, t2.y, t2.y, t2.z, -cdot);
      s0.x += df.x;
      s0.y += df.y;
      const V2 p = df.x * perp;
      const float la = df.x * s1_ + df.y;
      const float lb = df.x * s2_ + df.y;
      v_a = v_a - m_a * p;
      w_a -= i_a * la;
      v_b = v_b + m_b * p;
      w_b += i_b * lb;
    }
  } else if (jr.type == B2GPU_JOINT_WELD) {
    const float4 t4 = B.j_tmp[jt_at(B, x, j, 4)];
    const float m[9] = {t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w, t4.x};
    const float gamma = t4.y, bias = t4.z;
    if (jr.param[3] > 0.0f) {
      const float cdot2 = w_b - w_a;
      const float impulse2 = -m[8] * (cdot2 + bias + gamma * s0.z);
      s0.z += impulse2;
      w_a -= i_a * impulse2;
      w_b += i_b * impulse2;
      const V2 cdot1 = v_b + cross_sv(w_b, r_b) - v_a - cross_sv(w_a, r_a);
      const V2 impulse1 = -v2(m[0] * cdot1.x + m[3] * cdot1.y, m[1] * cdot1.x + m[4] * cdot1.y);  // b2_mul22
      s0.x += impulse1.x;
      s0.y += impulse1.y;
      const V2 p = impulse1;
      v_a = v_a - m_a * p;
      w_a -= i_a * cross(r_a, p);
      v_b = v_b + m_b * p;
      w_b += i_b * cross(r_b, p);
    } else {
      const V2 cdot1 = v_b + cross_sv(w_b, r_b) - v_a - cross_sv(w_a, r_a);
      const float cdot2 = w_b - w_a;
      // b2_mul_mat33 (src/b2_math.rs:601-603): v.x * ex + v.y * ey + v.z * ez, then the negation
      const float ix = -((cdot1.x * m[0] + cdot1.y * m[3]) + cdot2 * m[6]);
      const float iy = -((cdot1.x * m[1] + cdot1.y * m[4]) + cdot2 * m[7]);
      const float iz = -((cdot1.x * m[2] + cdot1.y * m[5]) + cdot2 * m[8]);
      s0.x += ix; s0.y += iy; s0.z += iz;
      const V2 p = v2(ix, iy);
      v_a = v_a - m_a * p;
      w_a -= i_a * (cross(r_a, p) + iz);
      v_b = v_b + m_b * p;
      w_b += i_b * (cross(r_b, p) + iz);
    }
  } else if (jr.type == B2GPU_JOINT_REVOLUTE) {
    const float axial_mass = t1.w, angle = t2.x;
    const bool fixed_rotation = i_a + i_b == 0.0f;
    if ((jflags & B2GPU_JOINT_ENABLE_MOTOR) && fixed_rotation == false) {
      const float cdot = w_b - w_a - s1.y;
      float impulse = -axial_mass * cdot;
      const float old_impulse = s0.z;
      const float max_impulse = h * s1.z;
      s0.z = fclamp_sel(s0.z + impulse, -max_impulse, max_impulse);
      impulse = s0.z - old_impulse;
      w_a -= i_a * impulse;
      w_b += i_b * impulse;
    }
    if ((jflags & B2GPU_JOINT_ENABLE_LIMIT) && fixed_rotation == false) {
      {  // lower limit
        const float c = angle - jr.param[1];
        const float cdot = w_b - w_a;
        float impulse = -axial_mass * (cdot + fmax_sel(c, 0.0f) * inv_dt);
        const float old_impulse = s0.w;
        s0.w = fmax_sel(s0.w + impulse, 0.0f);
        impulse = s0.w - old_impulse;
        w_a -= i_a * impulse;
        w_b += i_b * impulse;
      }
      {  // upper limit: signs flipped to keep c positive when the constraint is satisfied
        const float c = jr.param[2] - angle;
        const float cdot = w_a - w_b;
        float impulse = -axial_mass * (cdot + fmax_sel(c, 0.0f) * inv_dt);
        const float old_impulse = s1.x;
        s1.x = fmax_sel(s1.x + impulse, 0.0f);
        impulse = s1.x - old_impulse;
        w_a += i_a * impulse;
        w_b -= i_b * impulse;
      }
    }
    {  // point-to-point constraint
      const V2 cdot = v_b + cross_sv(w_b, r_b) - v_a - cross_sv(w_a, r_a);
      const V2 impulse = mat22_solve(t1.x, t1.y, t1.y, t1.z, -cdot);
      s0.x += impulse.x;
      s0.y += impulse.y;
      v_a = v_a - m_a * impulse;
      w_a -= i_a * cross(r_a, impulse);
      v_b = v_b + m_b * impulse;
      w_b += i_b * cross(r_b, impulse);
    }
  } else {  // distance
    const V2 u = v2(t1.x, t1.y);
    const float mass = t1.z, soft_mass = t1.w, gamma = t2.x, bias_ = t2.y, current_length = t2.z;
    const float min_length = jr.param[1], max_length = jr.param[2], stiffness = jr.param[3];
    if (min_length < max_length) {
      if (stiffness > 0.0f) {
        const V2 vp_a = v_a + cross_sv(w_a, r_a);
        const V2 vp_b = v_b + cross_sv(w_b, r_b);
        const float cdot = dot(u, vp_b - vp_a);
        const float impulse = -soft_mass * (cdot + bias_ + gamma * s0.x);
        s0.x += impulse;
        const V2 p = impulse * u;
        v_a = v_a - m_a * p;
        w_a -= i_a * cross(r_a, p);
        v_b = v_b + m_b * p;
        w_b += i_b * cross(r_b, p);
      }
      {  // lower
        const float c = current_length - min_length;
        const float bias = fmax_sel(0.0f, c) * inv_dt;
        const V2 vp_a = v_a + cross_sv(w_a, r_a);
        const V2 vp_b = v_b + cross_sv(w_b, r_b);
        const float cdot = dot(u, vp_b - vp_a);
        float impulse = -mass * (cdot + bias);
        const float old_impulse = s0.w;
        s0.w = fmax_sel(0.0f, s0.w + impulse);
        impulse = s0.w - old_impulse;
        const V2 p = impulse * u;
        v_a = v_a - m_a * p;
        w_a -= i_a * cross(r_a, p);
        v_b = v_b + m_b * p;
        w_b += i_b * cross(r_b, p);
      }
      {  // upper
        const float c = max_length - current_length;
        const float bias = fmax_sel(0.0f, c) * inv_dt;
        const V2 vp_a = v_a + cross_sv(w_a, r_a);
        const V2 vp_b = v_b + cross_sv(w_b, r_b);
        const float cdot = dot(u, vp_a - vp_b);
        float impulse = -mass * (cdot + bias);
        const float old_impulse = s1.x;
        s1.x = fmax_sel(0.0f, s1.x + impulse);
        impulse = s1.x - old_impulse;
        const V2 p = (-impulse) * u;
        v_a = v_a - m_a * p;
        w_a -= i_a * cross(r_a, p);
        v_b = v_b + m_b * p;
        w_b += i_b * cross(r_b, p);
      }
    } else {  // equal limits
      const V2 vp_a = v_a + cross_sv(w_a, r_a);
      const V2 vp_b = v_b + cross_sv(w_b, r_b);
      const float cdot = dot(u, vp_b - vp_a);
      const float impulse = -mass * cdot;
      s0.x += impulse;
      const V2 p = impulse * u;
      v_a = v_a - m_a * p;
      w_a -= i_a * cross(r_a, p);
      v_b = v_b + m_b * p;
      w_b += i_b * cross(r_b, p);
    }
  }
  B.j_s0[ji] = s0;
  B.j_s1[ji] = s1;
  if (m_a != 0.0f || i_a != 0.0f) st.set_vel(jr.body_a, make_float4(v_a.x, v_a.y, w_a, 0.0f));
  if (m_b != 0.0f || i_b != 0.0f) st.set_vel(jr.body_b, make_float4(v_b.x, v_b.y, w_b, 0.0f));
}

// solve_position_constraints of joint j; returns "within tolerance".  b_rot caches B2Rot::new of the running angle (the
// contact position solver relies on it), so it is refreshed whenever this joint changed an angle's bits.
template <class S>
B2G_HD bool joint_solve_position(const Batch& B, const WIdx& x, const S& st, int j) {
  const b2gpu_joint_rec& jr = B.joints[j];
  if (jr.type == B2GPU_JOINT_GEAR) return gear_solve_position(B, x, st, j);
  if (jr.type == B2GPU_JOINT_FRICTION || jr.type == B2GPU_JOINT_MOTOR || jr.type == B2GPU_JOINT_MOUSE) return true;  // no position rows
  const int bai = x.at(B.NB, jr.body_a), bbi = x.at(B.NB, jr.body_b);
  const float4 msa = B.b_mass[bai], msb = B.b_mass[bbi];
  const float m_a = msa.x, i_a = msa.y, m_b = msb.x, i_b = msb.y;
  float4 pa = st.pos(jr.body_a), pb = st.pos(jr.body_b);
  V2 c_a = v2(pa.x, pa.y), c_b = v2(pb.x, pb.y);
  float a_a = pa.z, a_b = pb.z;
  Rot q_a = st.rot(jr.body_a), q_b = st.rot(jr.body_b);
  const V2 la = v2(jr.local_anchor_a[0], jr.local_anchor_a[1]) - v2(msa.z, msa.w);
  const V2 lb = v2(jr.local_anchor_b[0], jr.local_anchor_b[1]) - v2(msb.z, msb.w);
  bool okay;
  if (jr.type == B2GPU_JOINT_WHEEL) {
    const int jflags = f2i(B.j_s1[x.at(B.NJ, j)].w);
    const float4 t0 = B.j_tmp[jt_at(B, x, j, 0)], t1 = B.j_tmp[jt_at(B, x, j, 1)];
    const V2 lx = v2(jr.param[5], jr.param[6]), ly = cross_sv(1.0f, lx);
    float linear_error = 0.0f;
    if (jflags & B2GPU_JOINT_ENABLE_LIMIT) {
      const V2 r_a = rot_mul(q_a, la);
      const V2 r_b = rot_mul(q_b, lb);
      const V2 d = (c_b - c_a) + r_b - r_a;
      const V2 ax = rot_mul(q_a, lx);
      const V2 ax0 = v2(t0.x, t0.y);  // m_ax of init_velocity_constraints: the reference crosses with it here
      const float s_ax = cross(d + r_a, ax0), s_bx = cross(r_b, ax0);
      float c = 0.0f;
      const float translation = dot(ax, d);
      const float lower = jr.param[1], upper = jr.param[2];
      if (fabsf(upper - lower) < 2.0f * B2G_LINEAR_SLOP) c = translation;
      else if (translation <= lower) c = fmin_sel(translation - lower, 0.0f);
      else if (translation >= upper) c = fmax_sel(translation - upper, 0.0f);
      if (c != 0.0f) {
        const float inv_mass = m_a + m_b + i_a * s_ax * s_ax + i_b * s_bx * s_bx;
        float impulse = 0.0f;
        if (inv_mass != 0.0f) impulse = -c / inv_mass;
        const V2 p = impulse * ax;
        const float la_ = impulse * s_ax, lb_ = impulse * s_bx;
        c_a = c_a - m_a * p;
        const float na = a_a - i_a * la_;
        c_b = c_b + m_b * p;
        const float nb = a_b + i_b * lb_;
        if (f2u(na) != f2u(a_a)) { a_a = na; q_a = rot_from_angle(na); }
        if (f2u(nb) != f2u(a_b)) { a_b = nb; q_b = rot_from_angle(nb); }
        linear_error = fabsf(c);
      }
    }
    {  // perpendicular constraint (rotations of the angles as they are now)
      const V2 r_a = rot_mul(q_a, la);
      const V2 r_b = rot_mul(q_b, lb);
      const V2 d = (c_b - c_a) + r_b - r_a;
      const V2 ay = rot_mul(q_a, ly);
      const float s_ay = cross(d + r_a, ay), s_by = cross(r_b, ay);
      const float c = dot(d, ay);
      // the reference uses m_s_ay / m_s_by of init_velocity_constraints in the effective mass
      const float inv_mass = m_a + m_b + i_a * t1.z * t1.z + i_b * t1.w * t1.w;
      float impulse = 0.0f;
      if (inv_mass != 0.0f) impulse = -c / inv_mass;
      const V2 p = impulse * ay;
      const float la_ = impulse * s_ay, lb_ = impulse * s_by;
      c_a = c_a - m_a * p;
      a_a -= i_a * la_;
      c_b = c_b + m_b * p;
      a_b += i_b * lb_;
      linear_error = fmax_sel(linear_error, fabsf(c));
    }
    okay = linear_error <= B2G_LINEAR_SLOP;
  } else if (jr.type == B2GPU_JOINT_PRISMATIC) {
    const int jflags = f2i(B.j_s1[x.at(B.NJ, j)].w);
    const V2 r_a = rot_mul(q_a, la);
    const V2 r_b = rot_mul(q_b, lb);
    const V2 d = c_b + r_b - c_a - r_a;
    const V2 lx = v2(jr.param[5], jr.param[6]), ly = cross_sv(1.0f, lx);
    const V2 axis = rot_mul(q_a, lx);
    const float a1 = cross(d + r_a, axis), a2 = cross(r_b, axis);
    const V2 perp = rot_mul(q_a, ly);
    const float s1 = cross(d + r_a, perp), s2 = cross(r_b, perp);
    const V2 c1 = v2(dot(perp, d), a_b - a_a - jr.param[0]);
    float linear_error = fabsf(c1.x);
    const float angular_error = fabsf(c1.y);
    bool active = false;
    float c2 = 0.0f;
    if (jflags & B2GPU_JOINT_ENABLE_LIMIT) {
      const float lower = jr.param[1], upper = jr.param[2];
      const float translation = dot(axis, d);
      if (fabsf(upper - lower) < 2.0f * B2G_LINEAR_SLOP) {
        c2 = translation;
        linear_error = fmax_sel(linear_error, fabsf(translation));
        active = true;
      } else if (translation <= lower) {
        c2 = fmin_sel(translation - lower, 0.0f);
        linear_error = fmax_sel(linear_error, lower - translation);
        active = true;
      } else if (translation >= upper) {
        c2 = fmax_sel(translation - upper, 0.0f);
        linear_error = fmax_sel(linear_error, translation - upper);
        active = true;
      }
    }
    float imp[3];
    const float k11 = m_a + m_b + i_a * s1 * s1 + i_b * s2 * s2;
    const float k12 = i_a * s1 + i_b * s2;
    float k22 = i_a + i_b;
    if (k22 == 0.0f) k22 = 1.0f;  // fixed rotation
    if (active) {
      const float k13 = i_a * s1 * a1 + i_b * s2 * a2;
      const float k23 = i_a * a1 + i_b * a2;
      const float k33 = m_a + m_b + i_a * a1 * a1 + i_b * a2 * a2;
      const float k[9] = {k11, k12, k13, k12, k22, k23, k13, k23, k33};
      const float c[3] = {-c1.x, -c1.y, -c2};
      mat33_solve33(k, c, imp);
    } else {
      const V2 impulse1 = mat22_solve(k11, k12, k12, k22, -c1);
      imp[0] = impulse1.x; imp[1] = impulse1.y; imp[2] = 0.0f;
    }
    const V2 p = imp[0] * perp + imp[2] * axis;
    const float la_ = imp[0] * s1 + imp[1] + imp[2] * a1;
    const float lb_ = imp[0] * s2 + imp[1] + imp[2] * a2;
    c_a = c_a - m_a * p;
    a_a -= i_a * la_;
    c_b = c_b + m_b * p;
    a_b += i_b * lb_;
    okay = linear_error <= B2G_LINEAR_SLOP && angular_error <= B2G_ANGULAR_SLOP;
  } else if (jr.type == B2GPU_JOINT_WELD) {
    const V2 r_a = rot_mul(q_a, la);
    const V2 r_b = rot_mul(q_b, lb);
    float k[9];
    weld_k(k, r_a, r_b, m_a, i_a, m_b, i_b);
    float position_error, angular_error;
    if (jr.param[3] > 0.0f) {
      const V2 c1 = c_b + r_b - c_a - r_a;
      position_error = length(c1);
      angular_error = 0.0f;
      const V2 p = -mat22_solve(k[0], k[3], k[1], k[4], c1);  // B2Mat33::solve22
      c_a = c_a - m_a * p;
      a_a -= i_a * cross(r_a, p);
      c_b = c_b + m_b * p;
      a_b += i_b * cross(r_b, p);
    } else {
      const V2 c1 = c_b + r_b - c_a - r_a;
      const float c2 = a_b - a_a - jr.param[0];
      position_error = length(c1);
      angular_error = fabsf(c2);
      float imp[3];
      if (k[8] > 0.0f) {
        const float c[3] = {c1.x, c1.y, c2};
        float xs[3];
        mat33_solve33(k, c, xs);
        imp[0] = -xs[0]; imp[1] = -xs[1]; imp[2] = -xs[2];
      } else {
        const V2 impulse2 = -mat22_solve(k[0], k[3], k[1], k[4], c1);
        imp[0] = impulse2.x; imp[1] = impulse2.y; imp[2] = 0.0f;
      }
      const V2 p = v2(imp[0], imp[1]);
      c_a = c_a - m_a * p;
      a_a -= i_a * (cross(r_a, p) + imp[2]);
      c_b = c_b + m_b * p;
      a_b += i_b * (cross(r_b, p) + imp[2]);
    }
    okay = position_error <= B2G_LINEAR_SLOP && angular_error <= B2G_ANGULAR_SLOP;
  } else if (jr.type == B2GPU_JOINT_REVOLUTE) {
    const float axial_mass = B.j_tmp[jt_at(B, x, j, 1)].w;
    const int jflags = f2i(B.j_s1[x.at(B.NJ, j)].w);
    float angular_error = 0.0f;
    const bool fixed_rotation = i_a + i_b == 0.0f;
    if ((jflags & B2GPU_JOINT_ENABLE_LIMIT) && fixed_rotation == false) {
      const float lower = jr.param[1], upper = jr.param[2];
      const float angle = a_b - a_a - jr.param[0];
      float c = 0.0f;
      if (fabsf(upper - lower) < 2.0f * B2G_ANGULAR_SLOP) {
        c = fclamp_sel(angle - lower, -B2G_MAX_ANGULAR_CORRECTION, B2G_MAX_ANGULAR_CORRECTION);
      } else if (angle <= lower) {
        c = fclamp_sel(angle - lower + B2G_ANGULAR_SLOP, -B2G_MAX_ANGULAR_CORRECTION, 0.0f);
      } else if (angle >= upper) {
        c = fclamp_sel(angle - upper - B2G_ANGULAR_SLOP, 0.0f, B2G_MAX_ANGULAR_CORRECTION);
      }
      const float limit_impulse = -axial_mass * c;
      const float na = a_a - i_a * limit_impulse, nb = a_b + i_b * limit_impulse;
      if (f2u(na) != f2u(a_a)) { a_a = na; q_a = rot_from_angle(na); }
      if (f2u(nb) != f2u(a_b)) { a_b = nb; q_b = rot_from_angle(nb); }
      angular_error = fabsf(c);
    }
    const V2 r_a = rot_mul(q_a, la);
    const V2 r_b = rot_mul(q_b, lb);
    const V2 c = c_b + r_b - c_a - r_a;
    const float position_error = length(c);
    const float kxx = m_a + m_b + i_a * r_a.y * r_a.y + i_b * r_b.y * r_b.y;
    const float kxy = -i_a * r_a.x * r_a.y - i_b * r_b.x * r_b.y;
    const float kyy = m_a + m_b + i_a * r_a.x * r_a.x + i_b * r_b.x * r_b.x;
    const V2 impulse = -mat22_solve(kxx, kxy, kxy, kyy, c);
    c_a = c_a - m_a * impulse;
    a_a -= i_a * cross(r_a, impulse);
    c_b = c_b + m_b * impulse;
    a_b += i_b * cross(r_b, impulse);
    okay = position_error <= B2G_LINEAR_SLOP && angular_error <= B2G_ANGULAR_SLOP;
  } else if (jr.type == B2GPU_JOINT_PULLEY) {  // private joints/b2_pulley_joint.rs:172-240
    const float ratio = jr.param[6];
    const V2 r_a = rot_mul(q_a, la);
    const V2 r_b = rot_mul(q_b, lb);
    V2 u_a = c_a + r_a - v2(jr.param[0], jr.param[1]);
    V2 u_b = c_b + r_b - v2(jr.param[2], jr.param[3]);
    const float length_a = length(u_a), length_b = length(u_b);
    if (length_a > 10.0f * B2G_LINEAR_SLOP) u_a = (1.0f / length_a) * u_a; else u_a = v2(0.0f, 0.0f);
    if (length_b > 10.0f * B2G_LINEAR_SLOP) u_b = (1.0f / length_b) * u_b; else u_b = v2(0.0f, 0.0f);
    const float ru_a = cross(r_a, u_a), ru_b = cross(r_b, u_b);
    const float pm_a = m_a + i_a * ru_a * ru_a;
    const float pm_b = m_b + i_b * ru_b * ru_b;
    float mass = pm_a + ratio * ratio * pm_b;
    if (mass > 0.0f) mass = 1.0f / mass;
    const float c = jr.param[7] - length_a - ratio * length_b;
    const float impulse = -mass * c;
    const V2 pa_ = (-impulse) * u_a;
    const V2 pb_ = (-ratio * impulse) * u_b;
    c_a = c_a + m_a * pa_;
    a_a += i_a * cross(r_a, pa_);
    c_b = c_b + m_b * pb_;
    a_b += i_b * cross(r_b, pb_);
    okay = fabsf(c) < B2G_LINEAR_SLOP;
  } else {  // distance
    const float mass = B.j_tmp[jt_at(B, x, j, 1)].z;
    const float min_length = jr.param[1], max_length = jr.param[2];
    const V2 r_a = rot_mul(q_a, la);
    const V2 r_b = rot_mul(q_b, lb);
    V2 u = c_b + r_b - c_a - r_a;
    const float len = normalize(u);
    float c;
    if (min_length == max_length) c = len - min_length;
    else if (len < min_length) c = len - min_length;
    else if (max_length < len) c = len - max_length;
    else return true;  // positions untouched
    const float impulse = -mass * c;
    const V2 p = impulse * u;
    c_a = c_a - m_a * p;
    a_a -= i_a * cross(r_a, p);
    c_b = c_b + m_b * p;
    a_b += i_b * cross(r_b, p);
    okay = fabsf(c) < B2G_LINEAR_SLOP;
  }
  if (m_a != 0.0f || i_a != 0.0f) {  // immovable bodies are shared between islands: never written
    if (f2u(a_a) != f2u(pa.z)) st.set_rot(jr.body_a, rot_from_angle(a_a));
    pa.x = c_a.x; pa.y = c_a.y; pa.z = a_a;
    st.set_pos(jr.body_a, pa);
  }
  if (m_b != 0.0f || i_b != 0.0f) {
    if (f2u(a_b) != f2u(pb.z)) st.set_rot(jr.body_b, rot_from_angle(a_b));
    pb.x = c_b.x; pb.y = c_b.y; pb.z = a_b;
    st.set_pos(jr.body_b, pb);
  }
  return okay;
}

}  // namespace b2g
