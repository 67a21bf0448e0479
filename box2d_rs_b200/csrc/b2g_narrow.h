// b2g_narrow.h — narrowphase manifold generators (one thread evaluates one contact).
//
// Reference: box2d-rs src/private/collision/b2_collide_circle.rs (:7 circles, :36 polygon/circle),
// b2_collide_polygon.rs (:8 find_max_separation, :48 find_incident_edge, :102 collide_polygons),
// b2_collide_edge.rs (:11 edge/circle, :172-228 axes, :230 edge/polygon),
// b2_collision.rs(private) (:7 world manifold, :171 clip_segment_to_line), and the shapes'
// compute_aabb (b2_circle_shape.rs:64, b2_edge_shape.rs:104, b2_polygon_shape.rs:292, private).
// Shapes are read through b2gpu_shape_rec (include/b2gpu.h): polygon vertices/normals as (x,y)
// pairs, an edge as v0..v3 in v[0..8), a circle centre in (cx,cy).
#pragma once
#include "../../include/b2gpu.h"
#include "b2g_math.h"

namespace b2g {

enum { FEAT_VERTEX = 0, FEAT_FACE = 1 };
B2G_HD uint32_t feat(int index_a, int index_b, int type_a, int type_b) {
  return (uint32_t)(index_a & 0xff) | ((uint32_t)(index_b & 0xff) << 8) | ((uint32_t)type_a << 16) | ((uint32_t)type_b << 24);
}
B2G_HD uint32_t feat_swap(uint32_t id) {  // exchange the A and B halves (manifold flip)
  return ((id >> 8) & 0xffu) | ((id & 0xffu) << 8) | (((id >> 24) & 0xffu) << 16) | (((id >> 16) & 0xffu) << 24);
}

struct Manifold {  // B2manifold, src/b2_collision.rs:104-114
  V2 pt[2];        // points[i].local_point
  float ni[2], ti[2];
  uint32_t id[2];
  V2 ln, lp;       // local_normal, local_point
  int type, count;
};
B2G_HD void manifold_clear(Manifold& m) {  // Default, src/b2_collision.rs:75-85
  for (int i = 0; i < 2; ++i) { m.pt[i] = v2(0.0f, 0.0f); m.ni[i] = 0.0f; m.ti[i] = 0.0f; m.id[i] = 0u; }
  m.ln = v2(0.0f, 0.0f);
  m.lp = v2(0.0f, 0.0f);
  m.type = B2GPU_MANIFOLD_CIRCLES;
  m.count = 0;
}

struct ClipV {
  V2 v;
  uint32_t id;
};

typedef const b2gpu_shape_rec* __restrict__ ShapeP;
B2G_HD V2 sh_vert(ShapeP s, int i) { return v2(s->v[2 * i], s->v[2 * i + 1]); }
B2G_HD V2 sh_norm(ShapeP s, int i) { return v2(s->n[2 * i], s->n[2 * i + 1]); }
B2G_HD V2 sh_center(ShapeP s) { return v2(s->cx, s->cy); }

// compute_aabb of one child shape under a transform.
B2G_HD Box shape_aabb(ShapeP s, const Xf& xf) {
  Box b;
  if (s->type == B2GPU_SHAPE_CIRCLE) {
    V2 p = xf.p + rot_mul(xf.q, sh_center(s));
    b.lo = v2(p.x - s->radius, p.y - s->radius);
    b.hi = v2(p.x + s->radius, p.y + s->radius);
  } else if (s->type == B2GPU_SHAPE_EDGE) {
    V2 v1 = xf_mul(xf, sh_vert(s, 1)), v2_ = xf_mul(xf, sh_vert(s, 2));
    V2 lo = vmin(v1, v2_), hi = vmax(v1, v2_);
    V2 r = v2(s->radius, s->radius);
    b.lo = lo - r;
    b.hi = hi + r;
  } else {
    V2 lo = xf_mul(xf, sh_vert(s, 0));
    V2 hi = lo;
    for (int i = 1; i < s->count; ++i) {
      V2 v = xf_mul(xf, sh_vert(s, i));
      lo = vmin(lo, v);
      hi = vmax(hi, v);
    }
    V2 r = v2(s->radius, s->radius);
    b.lo = lo - r;
    b.hi = hi + r;
  }
  return b;
}

// b2_clip_segment_to_line
B2G_HD int clip_segment(ClipV out[2], const ClipV in[2], V2 normal, float offset, int vertex_index_a) {
  int count = 0;
  float d0 = dot(normal, in[0].v) - offset;
  float d1 = dot(normal, in[1].v) - offset;
  if (d0 <= 0.0f) out[count++] = in[0];
  if (d1 <= 0.0f) out[count++] = in[1];
  if (d0 * d1 < 0.0f) {
    float interp = d0 / (d0 - d1);
    out[count].v = in[0].v + interp * (in[1].v - in[0].v);
    out[count].id = feat(vertex_index_a, (int)((in[0].id >> 8) & 0xffu), FEAT_VERTEX, FEAT_FACE);
    ++count;
  }
  return count;
}

B2G_HD void collide_circles(Manifold& m, ShapeP a, const Xf& xa, ShapeP b, const Xf& xb) {
  m.count = 0;
  V2 pa = xf_mul(xa, sh_center(a)), pb = xf_mul(xb, sh_center(b));
  V2 d = pb - pa;
  float dd = dot(d, d);
  float radius = a->radius + b->radius;
  if (dd > radius * radius) return;
  m.type = B2GPU_MANIFOLD_CIRCLES;
  m.lp = sh_center(a);
  m.ln = v2(0.0f, 0.0f);
  m.count = 1;
  m.pt[0] = sh_center(b);
  m.id[0] = 0u;
}

B2G_HD void collide_polygon_circle(Manifold& m, ShapeP poly, const Xf& xa, ShapeP circle, const Xf& xb) {
  m.count = 0;
  V2 c = xf_mul(xb, sh_center(circle));
  V2 cl = xf_mul_t(xa, c);
  int normal_index = 0;
  float separation = -B2G_MAX_FLOAT;
  float radius = poly->radius + circle->radius;
  int n = poly->count;
  for (int i = 0; i < n; ++i) {
    float s = dot(sh_norm(poly, i), cl - sh_vert(poly, i));
    if (s > radius) return;
    if (s > separation) { separation = s; normal_index = i; }
  }
  int i1 = normal_index;
  int i2 = i1 + 1 < n ? i1 + 1 : 0;
  V2 v1 = sh_vert(poly, i1), v2_ = sh_vert(poly, i2);
  // Fields are written only on the accepting paths: a rejected pair keeps the default manifold.
  V2 ln, lp;
  if (separation < B2G_EPSILON) {
    ln = sh_norm(poly, normal_index);
    lp = 0.5f * (v1 + v2_);
  } else {
    float u1 = dot(cl - v1, v2_ - v1);
    float u2 = dot(cl - v2_, v1 - v2_);
    if (u1 <= 0.0f) {
      if (dist_sq(cl, v1) > radius * radius) return;
      ln = cl - v1;
      normalize(ln);
      lp = v1;
    } else if (u2 <= 0.0f) {
      if (dist_sq(cl, v2_) > radius * radius) return;
      ln = cl - v2_;
      normalize(ln);
      lp = v2_;
    } else {
      V2 fc = 0.5f * (v1 + v2_);
      float s = dot(cl - fc, sh_norm(poly, i1));
      if (s > radius) return;
      ln = sh_norm(poly, i1);
      lp = fc;
    }
  }
  m.count = 1;
  m.type = B2GPU_MANIFOLD_FACE_A;
  m.ln = ln;
  m.lp = lp;
  m.pt[0] = sh_center(circle);
  m.id[0] = 0u;
}

B2G_HD float find_max_separation(int& edge_index, ShapeP p1, const Xf& xf1, ShapeP p2, const Xf& xf2) {
  int c1 = p1->count, c2 = p2->count;
  Xf xf = xf_mul_t_xf(xf2, xf1);
  int best = 0;
  float max_sep = -B2G_MAX_FLOAT;
  for (int i = 0; i < c1; ++i) {
    V2 n = rot_mul(xf.q, sh_norm(p1, i));
    V2 v1 = xf_mul(xf, sh_vert(p1, i));
    float si = B2G_MAX_FLOAT;
    for (int j = 0; j < c2; ++j) {
      float sij = dot(n, sh_vert(p2, j) - v1);
      if (sij < si) si = sij;
    }
    if (si > max_sep) { max_sep = si; best = i; }
  }
  edge_index = best;
  return max_sep;
}

B2G_HD void find_incident_edge(ClipV c[2], ShapeP p1, const Xf& xf1, int edge1, ShapeP p2, const Xf& xf2) {
  int c2 = p2->count;
  V2 normal1 = rot_mul_t(xf2.q, rot_mul(xf1.q, sh_norm(p1, edge1)));
  int index = 0;
  float min_dot = B2G_MAX_FLOAT;
  for (int i = 0; i < c2; ++i) {
    float d = dot(normal1, sh_norm(p2, i));
    if (d < min_dot) { min_dot = d; index = i; }
  }
  int i1 = index;
  int i2 = i1 + 1 < c2 ? i1 + 1 : 0;
  c[0].v = xf_mul(xf2, sh_vert(p2, i1));
  c[0].id = feat(edge1, i1, FEAT_FACE, FEAT_VERTEX);
  c[1].v = xf_mul(xf2, sh_vert(p2, i2));
  c[1].id = feat(edge1, i2, FEAT_FACE, FEAT_VERTEX);
}

B2G_HD void collide_polygons(Manifold& m, ShapeP pa, const Xf& xa, ShapeP pb, const Xf& xb) {
  m.count = 0;
  float total_radius = pa->radius + pb->radius;
  int edge_a = 0;
  float sep_a = find_max_separation(edge_a, pa, xa, pb, xb);
  if (sep_a > total_radius) return;
  int edge_b = 0;
  float sep_b = find_max_separation(edge_b, pb, xb, pa, xa);
  if (sep_b > total_radius) return;
  ShapeP p1;
  ShapeP p2;
  Xf xf1, xf2;
  int edge1;
  bool flip;
  const float k_tol = 0.1f * B2G_LINEAR_SLOP;
  if (sep_b > sep_a + k_tol) {
    p1 = pb; p2 = pa; xf1 = xb; xf2 = xa; edge1 = edge_b;
    m.type = B2GPU_MANIFOLD_FACE_B;
    flip = true;
  } else {
    p1 = pa; p2 = pb; xf1 = xa; xf2 = xb; edge1 = edge_a;
    m.type = B2GPU_MANIFOLD_FACE_A;
    flip = false;
  }
  ClipV incident[2];
  find_incident_edge(incident, p1, xf1, edge1, p2, xf2);
  int count1 = p1->count;
  int iv1 = edge1;
  int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  V2 v11 = sh_vert(p1, iv1), v12 = sh_vert(p1, iv2);
  V2 local_tangent = v12 - v11;
  normalize(local_tangent);
  V2 local_normal = cross_vs(local_tangent, 1.0f);
  V2 plane_point = 0.5f * (v11 + v12);
  V2 tangent = rot_mul(xf1.q, local_tangent);
  V2 normal = cross_vs(tangent, 1.0f);
  v11 = xf_mul(xf1, v11);
  v12 = xf_mul(xf1, v12);
  float front_offset = dot(normal, v11);
  float side_offset1 = -dot(tangent, v11) + total_radius;
  float side_offset2 = dot(tangent, v12) + total_radius;
  ClipV cp1[2], cp2[2];
  int np = clip_segment(cp1, incident, -tangent, side_offset1, iv1);
  if (np < 2) return;
  np = clip_segment(cp2, cp1, tangent, side_offset2, iv2);
  if (np < 2) return;
  m.ln = local_normal;
  m.lp = plane_point;
  int pc = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = dot(normal, cp2[i].v) - front_offset;
    if (separation <= total_radius) {
      m.pt[pc] = xf_mul_t(xf2, cp2[i].v);
      m.id[pc] = flip ? feat_swap(cp2[i].id) : cp2[i].id;
      ++pc;
    }
  }
  m.count = pc;
}

B2G_HD void collide_edge_circle(Manifold& m, ShapeP edge, const Xf& xa, ShapeP circle, const Xf& xb) {
  m.count = 0;
  V2 q = xf_mul_t(xa, xf_mul(xb, sh_center(circle)));
  V2 a = sh_vert(edge, 1), b = sh_vert(edge, 2);
  V2 e = b - a;
  V2 n = v2(e.y, -e.x);
  float offset = dot(n, q - a);
  bool one_sided = edge->one_sided != 0;
  if (one_sided && offset < 0.0f) return;
  float u = dot(e, b - q);
  float v = dot(e, q - a);
  float radius = edge->radius + circle->radius;
  if (v <= 0.0f) {
    V2 d = q - a;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (one_sided) {
      V2 a1 = sh_vert(edge, 0);
      V2 e1 = a - a1;
      float u1 = dot(e1, a - q);
      if (u1 > 0.0f) return;
    }
    m.count = 1;
    m.type = B2GPU_MANIFOLD_CIRCLES;
    m.ln = v2(0.0f, 0.0f);
    m.lp = a;
    m.id[0] = feat(0, 0, FEAT_VERTEX, FEAT_VERTEX);
    m.pt[0] = sh_center(circle);
    return;
  }
  if (u <= 0.0f) {
    V2 d = q - b;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (one_sided) {
      V2 b2 = sh_vert(edge, 3);
      V2 e2 = b2 - b;
      float v2q = dot(e2, q - b);
      if (v2q > 0.0f) return;
    }
    m.count = 1;
    m.type = B2GPU_MANIFOLD_CIRCLES;
    m.ln = v2(0.0f, 0.0f);
    m.lp = b;
    m.id[0] = feat(1, 0, FEAT_VERTEX, FEAT_VERTEX);
    m.pt[0] = sh_center(circle);
    return;
  }
  float den = dot(e, e);
  V2 p = (1.0f / den) * (u * a + v * b);
  V2 d = q - p;
  float dd = dot(d, d);
  if (dd > radius * radius) return;
  if (offset < 0.0f) n = v2(-n.x, -n.y);
  normalize(n);
  m.count = 1;
  m.type = B2GPU_MANIFOLD_FACE_A;
  m.ln = n;
  m.lp = a;
  m.id[0] = feat(0, 0, FEAT_FACE, FEAT_VERTEX);
  m.pt[0] = sh_center(circle);
}

enum { AXIS_UNKNOWN = 0, AXIS_EDGE_A = 1, AXIS_EDGE_B = 2 };
struct EPAxis {
  V2 normal;
  int type, index;
  float separation;
};

B2G_HD void collide_edge_polygon(Manifold& m, ShapeP edge, const Xf& xa, ShapeP pb, const Xf& xb) {
  m.count = 0;
  Xf xf = xf_mul_t_xf(xa, xb);
  V2 centroid_b = xf_mul(xf, sh_center(pb));
  V2 v1 = sh_vert(edge, 1), v2_ = sh_vert(edge, 2);
  V2 edge1 = v2_ - v1;
  normalize(edge1);
  V2 normal1 = v2(edge1.y, -edge1.x);
  float offset1 = dot(normal1, centroid_b - v1);
  bool one_sided = edge->one_sided != 0;
  if (one_sided && offset1 < 0.0f) return;
  // polygon B in frame A (b2_collide_edge.rs:262-269)
  V2 tv[B2G_MAX_POLY], tn[B2G_MAX_POLY];
  int tcount = pb->count;
  for (int i = 0; i < tcount; ++i) {
    tv[i] = xf_mul(xf, sh_vert(pb, i));
    tn[i] = rot_mul(xf.q, sh_norm(pb, i));
  }
  float radius = pb->radius + edge->radius;
  // b2_compute_edge_separation (:172-202)
  EPAxis edge_axis;
  edge_axis.type = AXIS_EDGE_A;
  edge_axis.index = -1;
  edge_axis.separation = -B2G_MAX_FLOAT;
  edge_axis.normal = v2(0.0f, 0.0f);
  for (int j = 0; j < 2; ++j) {
    V2 ax = j == 0 ? normal1 : -normal1;
    float sj = B2G_MAX_FLOAT;
    for (int i = 0; i < tcount; ++i) {
      float si = dot(ax, tv[i] - v1);
      if (si < sj) sj = si;
    }
    if (sj > edge_axis.separation) { edge_axis.index = j; edge_axis.separation = sj; edge_axis.normal = ax; }
  }
  if (edge_axis.separation > radius) return;
  // b2_compute_polygon_separation (:204-228)
  EPAxis poly_axis;
  poly_axis.type = AXIS_UNKNOWN;
  poly_axis.index = -1;
  poly_axis.separation = -B2G_MAX_FLOAT;
  poly_axis.normal = v2(0.0f, 0.0f);
  for (int i = 0; i < tcount; ++i) {
    V2 n = -tn[i];
    float s1 = dot(n, tv[i] - v1);
    float s2 = dot(n, tv[i] - v2_);
    float s = fmin_sel(s1, s2);
    if (s > poly_axis.separation) { poly_axis.type = AXIS_EDGE_B; poly_axis.index = i; poly_axis.separation = s; poly_axis.normal = n; }
  }
  if (poly_axis.separation > radius) return;
  const float k_relative_tol = 0.98f, k_absolute_tol = 0.001f;
  EPAxis primary;
  if (poly_axis.separation - radius > k_relative_tol * (edge_axis.separation - radius) + k_absolute_tol) primary = poly_axis;
  else primary = edge_axis;
  if (one_sided) {
    V2 edge0 = v1 - sh_vert(edge, 0);
    normalize(edge0);
    V2 normal0 = v2(edge0.y, -edge0.x);
    bool convex1 = cross(edge0, edge1) >= 0.0f;
    V2 edge2 = sh_vert(edge, 3) - v2_;
    normalize(edge2);
    V2 normal2 = v2(edge2.y, -edge2.x);
    bool convex2 = cross(edge1, edge2) >= 0.0f;
    const float sin_tol = 0.1f;
    bool side1 = dot(primary.normal, edge1) <= 0.0f;
    if (side1) {
      if (convex1) {
        if (cross(primary.normal, normal0) > sin_tol) return;
      } else {
        primary = edge_axis;
      }
    } else {
      if (convex2) {
        if (cross(normal2, primary.normal) > sin_tol) return;
      } else {
        primary = edge_axis;
      }
    }
  }
  ClipV cp[2];
  int rf_i1, rf_i2;
  V2 rf_v1, rf_v2, rf_normal, rf_sn1, rf_sn2;
  if (primary.type == AXIS_EDGE_A) {
    m.type = B2GPU_MANIFOLD_FACE_A;
    int best = 0;
    float best_value = dot(primary.normal, tn[0]);
    for (int i = 1; i < tcount; ++i) {
      float value = dot(primary.normal, tn[i]);
      if (value < best_value) { best_value = value; best = i; }
    }
    int i1 = best;
    int i2 = i1 + 1 < tcount ? i1 + 1 : 0;
    cp[0].v = tv[i1];
    cp[0].id = feat(0, i1, FEAT_FACE, FEAT_VERTEX);
    cp[1].v = tv[i2];
    cp[1].id = feat(0, i2, FEAT_FACE, FEAT_VERTEX);
    rf_i1 = 0;
    rf_i2 = 1;
    rf_v1 = v1;
    rf_v2 = v2_;
    rf_normal = primary.normal;
    rf_sn1 = -edge1;
    rf_sn2 = edge1;
  } else {
    m.type = B2GPU_MANIFOLD_FACE_B;
    cp[0].v = v2_;
    cp[0].id = feat(1, primary.index, FEAT_VERTEX, FEAT_FACE);
    cp[1].v = v1;
    cp[1].id = feat(0, primary.index, FEAT_VERTEX, FEAT_FACE);
    rf_i1 = primary.index;
    rf_i2 = rf_i1 + 1 < tcount ? rf_i1 + 1 : 0;
    rf_v1 = tv[rf_i1];
    rf_v2 = tv[rf_i2];
    rf_normal = tn[rf_i1];
    rf_sn1 = v2(rf_normal.y, -rf_normal.x);
    rf_sn2 = -rf_sn1;
  }
  float rf_so1 = dot(rf_sn1, rf_v1);
  float rf_so2 = dot(rf_sn2, rf_v2);
  ClipV cp1[2], cp2[2];
  int np = clip_segment(cp1, cp, rf_sn1, rf_so1, rf_i1);
  if (np < 2) return;
  np = clip_segment(cp2, cp1, rf_sn2, rf_so2, rf_i2);
  if (np < 2) return;
  if (primary.type == AXIS_EDGE_A) {
    m.ln = rf_normal;
    m.lp = rf_v1;
  } else {
    m.ln = sh_norm(pb, rf_i1);
    m.lp = sh_vert(pb, rf_i1);
  }
  int pc = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = dot(rf_normal, cp2[i].v - rf_v1);
    if (separation <= radius) {
      if (primary.type == AXIS_EDGE_A) {
        m.pt[pc] = xf_mul_t(xf, cp2[i].v);
        m.id[pc] = cp2[i].id;
      } else {
        m.pt[pc] = cp2[i].v;
        m.id[pc] = feat_swap(cp2[i].id);
      }
      ++pc;
    }
  }
  m.count = pc;
}

// Typed contact dispatch (contacts/*.rs:45-56).  Fixture A is always the "larger" type
// (b2_contact.rs(private):26-30), chains arrive as their child edge record.
B2G_HD void evaluate_contact(Manifold& m, ShapeP sa, const Xf& xa, ShapeP sb, const Xf& xb) {
  int ta = sa->type, tb = sb->type;
  if (ta == B2GPU_SHAPE_POLYGON && tb == B2GPU_SHAPE_POLYGON) collide_polygons(m, sa, xa, sb, xb);
  else if (ta == B2GPU_SHAPE_POLYGON && tb == B2GPU_SHAPE_CIRCLE) collide_polygon_circle(m, sa, xa, sb, xb);
  else if (ta == B2GPU_SHAPE_CIRCLE && tb == B2GPU_SHAPE_CIRCLE) collide_circles(m, sa, xa, sb, xb);
  else if (ta == B2GPU_SHAPE_EDGE && tb == B2GPU_SHAPE_POLYGON) collide_edge_polygon(m, sa, xa, sb, xb);
  else if (ta == B2GPU_SHAPE_EDGE && tb == B2GPU_SHAPE_CIRCLE) collide_edge_circle(m, sa, xa, sb, xb);
  else m.count = 0;
}

// B2worldManifold::initialize (b2_collision.rs(private):7-68) — normal and points only.
B2G_HD void world_manifold(V2& normal, V2 points[2], const Manifold& m, const Xf& xa, float ra, const Xf& xb, float rb) {
  normal = v2(0.0f, 0.0f);
  points[0] = points[1] = v2(0.0f, 0.0f);
  if (m.count == 0) return;
  if (m.type == B2GPU_MANIFOLD_CIRCLES) {
    normal = v2(1.0f, 0.0f);
    V2 pa = xf_mul(xa, m.lp);
    V2 pb = xf_mul(xb, m.pt[0]);
    if (dist_sq(pa, pb) > B2G_EPSILON * B2G_EPSILON) {
      normal = pb - pa;
      normalize(normal);
    }
    V2 ca = pa + ra * normal;
    V2 cb = pb - rb * normal;
    points[0] = 0.5f * (ca + cb);
  } else if (m.type == B2GPU_MANIFOLD_FACE_A) {
    normal = rot_mul(xa.q, m.ln);
    V2 plane = xf_mul(xa, m.lp);
    for (int i = 0; i < m.count; ++i) {
      V2 clip = xf_mul(xb, m.pt[i]);
      V2 ca = clip + (ra - dot(clip - plane, normal)) * normal;
      V2 cb = clip - rb * normal;
      points[i] = 0.5f * (ca + cb);
    }
  } else {
    normal = rot_mul(xb.q, m.ln);
    V2 plane = xf_mul(xb, m.lp);
    for (int i = 0; i < m.count; ++i) {
      V2 clip = xf_mul(xa, m.pt[i]);
      V2 cb = clip + (rb - dot(clip - plane, normal)) * normal;
      V2 ca = clip - ra * normal;
      points[i] = 0.5f * (ca + cb);
    }
    normal = -normal;
  }
}

}  // namespace b2g
