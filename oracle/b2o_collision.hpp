// TEST INFRASTRUCTURE — CPU oracle (see b2o_math.hpp header). PARITY UNPINNED beyond the
// reference's own tests.
//
// b2o_collision.hpp — restates src/b2_collision.rs, src/private/collision/b2_collision.rs,
// b2_collide_{circle,polygon,edge}.rs and the shape files' set/compute_aabb/compute_mass.
#pragma once
#include <cassert>
#include <vector>

#include "b2o_math.hpp"

namespace b2o {

enum ShapeType { E_CIRCLE = 0, E_EDGE = 1, E_POLYGON = 2, E_CHAIN = 3 };
enum ManifoldType { E_CIRCLES = 0, E_FACE_A = 1, E_FACE_B = 2 };
enum FeatureType { E_VERTEX = 0, E_FACE = 1 };

// src/b2_collision.rs:23-32
struct ContactFeature {
  uint8_t index_a = 0, index_b = 0, type_a = 0, type_b = 0;
  bool operator==(const ContactFeature& o) const {
    return index_a == o.index_a && index_b == o.index_b && type_a == o.type_a && type_b == o.type_b;
  }
  uint32_t key() const { return index_a | (index_b << 8) | (type_a << 16) | (type_b << 24); }
};
struct ManifoldPoint {
  Vec2 local_point;
  float normal_impulse = 0.0f, tangent_impulse = 0.0f;
  ContactFeature id;
};
struct Manifold {  // :104-114, Default :75-85
  ManifoldPoint points[MAX_MANIFOLD_POINTS];
  Vec2 local_normal, local_point;
  int type = E_CIRCLES;
  int point_count = 0;
};
struct WorldManifold {
  Vec2 normal;
  Vec2 points[MAX_MANIFOLD_POINTS];
  float separations[MAX_MANIFOLD_POINTS] = {0.0f, 0.0f};
};
struct ClipVertex {
  Vec2 v;
  ContactFeature id;
};
struct MassData {
  float mass = 0.0f;
  Vec2 center;
  float i = 0.0f;
};

// One shape object as a fixture owns it (clone of the user's shape).
struct Shape {
  int type = E_CIRCLE;
  float radius = 0.0f;
  // circle
  Vec2 p;
  // edge
  Vec2 v0, v1, v2, v3;
  bool one_sided = false;
  // polygon
  Vec2 centroid;
  Vec2 vertices[MAX_POLYGON_VERTICES];
  Vec2 normals[MAX_POLYGON_VERTICES];
  int count = 0;
  // chain
  std::vector<Vec2> chain;
  Vec2 chain_prev, chain_next;

  int child_count() const { return type == E_CHAIN ? (int)chain.size() - 1 : 1; }
};

// ---- polygon: b2_polygon_shape.rs(private)
inline void polygon_set_as_box(Shape& s, float hx, float hy) {  // :15-27
  s.type = E_POLYGON;
  s.radius = POLYGON_RADIUS;
  s.count = 4;
  s.vertices[0].set(-hx, -hy);
  s.vertices[1].set(hx, -hy);
  s.vertices[2].set(hx, hy);
  s.vertices[3].set(-hx, hy);
  s.normals[0].set(0.0f, -1.0f);
  s.normals[1].set(1.0f, 0.0f);
  s.normals[2].set(0.0f, 1.0f);
  s.normals[3].set(-1.0f, 0.0f);
  s.centroid.set_zero();
}
inline void polygon_set_as_box_angle(Shape& s, float hx, float hy, Vec2 center, float angle) {  // :29-57
  polygon_set_as_box(s, hx, hy);
  s.centroid = center;
  Transform xf;
  xf.p = center;
  xf.q = Rot(angle);
  for (int i = 0; i < s.count; ++i) {
    s.vertices[i] = b2_mul_xf(xf, s.vertices[i]);
    s.normals[i] = b2_mul_rot(xf.q, s.normals[i]);
  }
}
inline Vec2 compute_centroid(const Vec2* vs, int count) {  // :63-94
  Vec2 c(0.0f, 0.0f);
  float area = 0.0f;
  Vec2 s = vs[0];
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < count; ++i) {
    Vec2 p1 = vs[0] - s;
    Vec2 p2 = vs[i] - s;
    Vec2 p3 = i + 1 < count ? vs[i + 1] - s : vs[0] - s;
    Vec2 e1 = p2 - p1, e2 = p3 - p1;
    float d = b2_cross(e1, e2);
    float triangle_area = 0.5f * d;
    area += triangle_area;
    c += (triangle_area * inv3) * (p1 + p2 + p3);
  }
  assert(area > EPSILON);
  c = (1.0f / area) * c + s;
  return c;
}
inline bool polygon_set(Shape& s, const Vec2* vertices, int count) {  // :96-211
  s.type = E_POLYGON;
  s.radius = POLYGON_RADIUS;
  if (count < 3) { polygon_set_as_box(s, 1.0f, 1.0f); return false; }
  int n = b2_min(count, MAX_POLYGON_VERTICES);
  Vec2 ps[MAX_POLYGON_VERTICES];
  int temp_count = 0;
  for (int i = 0; i < n; ++i) {
    Vec2 v = vertices[i];
    bool unique = true;
    for (int j = 0; j < temp_count; ++j) {
      if (b2_distance_squared(v, ps[j]) < ((0.5f * LINEAR_SLOP) * (0.5f * LINEAR_SLOP))) { unique = false; break; }
    }
    if (unique) ps[temp_count++] = v;
  }
  n = temp_count;
  if (n < 3) { polygon_set_as_box(s, 1.0f, 1.0f); return false; }
  int i0 = 0;
  float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
  }
  int hull[MAX_POLYGON_VERTICES];
  int m = 0, ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      Vec2 r = ps[ie] - ps[hull[m]];
      Vec2 v = ps[j] - ps[hull[m]];
      float c = b2_cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && v.length_squared() > r.length_squared()) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  if (m < 3) { polygon_set_as_box(s, 1.0f, 1.0f); return false; }
  s.count = m;
  for (int i = 0; i < m; ++i) s.vertices[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i2 = i + 1 < m ? i + 1 : 0;
    Vec2 edge = s.vertices[i2] - s.vertices[i];
    s.normals[i] = b2_cross_vs(edge, 1.0f);
    s.normals[i].normalize();
  }
  s.centroid = compute_centroid(s.vertices, m);
  return true;
}

// ---- compute_mass: circle :79-86, edge :126-132, polygon :314-390, chain :135-141
inline void shape_compute_mass(const Shape& s, MassData& md, float density) {
  switch (s.type) {
    case E_CIRCLE:
      md.mass = density * PI * s.radius * s.radius;
      md.center = s.p;
      md.i = md.mass * (0.5f * s.radius * s.radius + b2_dot(s.p, s.p));
      break;
    case E_EDGE:
      md.mass = 0.0f;
      md.center = 0.5f * (s.v1 + s.v2);
      md.i = 0.0f;
      break;
    case E_CHAIN:
      md.mass = 0.0f;
      md.center.set_zero();
      md.i = 0.0f;
      break;
    case E_POLYGON: {
      Vec2 center(0.0f, 0.0f);
      float area = 0.0f, inert = 0.0f;
      Vec2 ref = s.vertices[0];
      const float k_inv3 = 1.0f / 3.0f;
      for (int i = 0; i < s.count; ++i) {
        Vec2 e1 = s.vertices[i] - ref;
        Vec2 e2 = i + 1 < s.count ? s.vertices[i + 1] - ref : s.vertices[0] - ref;
        float d = b2_cross(e1, e2);
        float triangle_area = 0.5f * d;
        area += triangle_area;
        center += (triangle_area * k_inv3) * (e1 + e2);
        float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
        float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
        float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
        inert += (0.25f * k_inv3 * d) * (intx2 + inty2);
      }
      md.mass = density * area;
      assert(area > EPSILON);
      center *= 1.0f / area;
      md.center = center + ref;
      md.i = density * inert;
      md.i += md.mass * (b2_dot(md.center, md.center) - b2_dot(center, center));
    } break;
  }
}

// b2_chain_shape.rs(private):59-79
inline void chain_get_child_edge(const Shape& chain, Shape& edge, int index) {
  edge.type = E_EDGE;
  edge.radius = chain.radius;
  edge.v1 = chain.chain[index + 0];
  edge.v2 = chain.chain[index + 1];
  edge.one_sided = true;
  edge.v0 = index > 0 ? chain.chain[index - 1] : chain.chain_prev;
  edge.v3 = index < (int)chain.chain.size() - 2 ? chain.chain[index + 2] : chain.chain_next;
}

// ---- compute_aabb: circle :64-77, edge :104-124, polygon :292-312, chain :109-133
inline void shape_compute_aabb(const Shape& s, AABB& aabb, const Transform& xf, int child) {
  switch (s.type) {
    case E_CIRCLE: {
      Vec2 p = xf.p + b2_mul_rot(xf.q, s.p);
      aabb.lower.set(p.x - s.radius, p.y - s.radius);
      aabb.upper.set(p.x + s.radius, p.y + s.radius);
    } break;
    case E_EDGE:
    case E_CHAIN: {
      Vec2 a, b;
      if (s.type == E_EDGE) { a = s.v1; b = s.v2; }
      else {
        int i1 = child, i2 = child + 1;
        if (i2 == (int)s.chain.size()) i2 = 0;
        a = s.chain[i1]; b = s.chain[i2];
      }
      Vec2 v1 = b2_mul_xf(xf, a), v2 = b2_mul_xf(xf, b);
      Vec2 lower = b2_min_v(v1, v2), upper = b2_max_v(v1, v2);
      Vec2 r(s.radius, s.radius);
      aabb.lower = lower - r;
      aabb.upper = upper + r;
    } break;
    case E_POLYGON: {
      Vec2 lower = b2_mul_xf(xf, s.vertices[0]);
      Vec2 upper = lower;
      for (int i = 1; i < s.count; ++i) {
        Vec2 v = b2_mul_xf(xf, s.vertices[i]);
        lower = b2_min_v(lower, v);
        upper = b2_max_v(upper, v);
      }
      Vec2 r(s.radius, s.radius);
      aabb.lower = lower - r;
      aabb.upper = upper + r;
    } break;
  }
}

// ---- b2_collision.rs(private):7-68
inline void world_manifold_initialize(WorldManifold& wm, const Manifold& m, const Transform& xf_a, float radius_a,
                                      const Transform& xf_b, float radius_b) {
  if (m.point_count == 0) return;
  switch (m.type) {
    case E_CIRCLES: {
      wm.normal.set(1.0f, 0.0f);
      Vec2 point_a = b2_mul_xf(xf_a, m.local_point);
      Vec2 point_b = b2_mul_xf(xf_b, m.points[0].local_point);
      if (b2_distance_squared(point_a, point_b) > EPSILON * EPSILON) {
        wm.normal = point_b - point_a;
        wm.normal.normalize();
      }
      Vec2 c_a = point_a + radius_a * wm.normal;
      Vec2 c_b = point_b - radius_b * wm.normal;
      wm.points[0] = 0.5f * (c_a + c_b);
      wm.separations[0] = b2_dot(c_b - c_a, wm.normal);
    } break;
    case E_FACE_A: {
      wm.normal = b2_mul_rot(xf_a.q, m.local_normal);
      Vec2 plane_point = b2_mul_xf(xf_a, m.local_point);
      for (int i = 0; i < m.point_count; ++i) {
        Vec2 clip_point = b2_mul_xf(xf_b, m.points[i].local_point);
        Vec2 c_a = clip_point + (radius_a - b2_dot(clip_point - plane_point, wm.normal)) * wm.normal;
        Vec2 c_b = clip_point - radius_b * wm.normal;
        wm.points[i] = 0.5f * (c_a + c_b);
        wm.separations[i] = b2_dot(c_b - c_a, wm.normal);
      }
    } break;
    case E_FACE_B: {
      wm.normal = b2_mul_rot(xf_b.q, m.local_normal);
      Vec2 plane_point = b2_mul_xf(xf_b, m.local_point);
      for (int i = 0; i < m.point_count; ++i) {
        Vec2 clip_point = b2_mul_xf(xf_a, m.points[i].local_point);
        Vec2 c_b = clip_point + (radius_b - b2_dot(clip_point - plane_point, wm.normal)) * wm.normal;
        Vec2 c_a = clip_point - radius_a * wm.normal;
        wm.points[i] = 0.5f * (c_a + c_b);
        wm.separations[i] = b2_dot(c_a - c_b, wm.normal);
      }
      wm.normal = -wm.normal;
    } break;
  }
}

// ---- b2_collision.rs(private):171-214
inline int clip_segment_to_line(ClipVertex v_out[2], const ClipVertex v_in[2], Vec2 normal, float offset,
                                int vertex_index_a) {
  int count = 0;
  float distance0 = b2_dot(normal, v_in[0].v) - offset;
  float distance1 = b2_dot(normal, v_in[1].v) - offset;
  if (distance0 <= 0.0f) v_out[count++] = v_in[0];
  if (distance1 <= 0.0f) v_out[count++] = v_in[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    v_out[count].v = v_in[0].v + interp * (v_in[1].v - v_in[0].v);
    v_out[count].id.index_a = (uint8_t)vertex_index_a;
    v_out[count].id.index_b = v_in[0].id.index_b;
    v_out[count].id.type_a = E_VERTEX;
    v_out[count].id.type_b = E_FACE;
    ++count;
  }
  return count;
}

// ---- b2_collide_circle.rs:7-34
inline void collide_circles(Manifold& m, const Shape& a, const Transform& xf_a, const Shape& b, const Transform& xf_b) {
  m.point_count = 0;
  Vec2 p_a = b2_mul_xf(xf_a, a.p), p_b = b2_mul_xf(xf_b, b.p);
  Vec2 d = p_b - p_a;
  float dist_sqr = b2_dot(d, d);
  float radius = a.radius + b.radius;
  if (dist_sqr > radius * radius) return;
  m.type = E_CIRCLES;
  m.local_point = a.p;
  m.local_normal.set_zero();
  m.point_count = 1;
  m.points[0].local_point = b.p;
  m.points[0].id = ContactFeature();
}

// ---- b2_collide_circle.rs:36-133
inline void collide_polygon_and_circle(Manifold& m, const Shape& poly, const Transform& xf_a, const Shape& circle,
                                       const Transform& xf_b) {
  m.point_count = 0;
  Vec2 c = b2_mul_xf(xf_b, circle.p);
  Vec2 c_local = b2_mul_t_xf(xf_a, c);
  int normal_index = 0;
  float separation = -MAX_FLOAT;
  float radius = poly.radius + circle.radius;
  int vertex_count = poly.count;
  for (int i = 0; i < vertex_count; ++i) {
    float s = b2_dot(poly.normals[i], c_local - poly.vertices[i]);
    if (s > radius) return;
    if (s > separation) { separation = s; normal_index = i; }
  }
  int vi1 = normal_index;
  int vi2 = vi1 + 1 < vertex_count ? vi1 + 1 : 0;
  Vec2 v1 = poly.vertices[vi1], v2 = poly.vertices[vi2];
  if (separation < EPSILON) {
    m.point_count = 1;
    m.type = E_FACE_A;
    m.local_normal = poly.normals[normal_index];
    m.local_point = 0.5f * (v1 + v2);
    m.points[0].local_point = circle.p;
    m.points[0].id = ContactFeature();
    return;
  }
  float u1 = b2_dot(c_local - v1, v2 - v1);
  float u2 = b2_dot(c_local - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (b2_distance_squared(c_local, v1) > radius * radius) return;
    m.point_count = 1;
    m.type = E_FACE_A;
    m.local_normal = c_local - v1;
    m.local_normal.normalize();
    m.local_point = v1;
    m.points[0].local_point = circle.p;
    m.points[0].id = ContactFeature();
  } else if (u2 <= 0.0f) {
    if (b2_distance_squared(c_local, v2) > radius * radius) return;
    m.point_count = 1;
    m.type = E_FACE_A;
    m.local_normal = c_local - v2;
    m.local_normal.normalize();
    m.local_point = v2;
    m.points[0].local_point = circle.p;
    m.points[0].id = ContactFeature();
  } else {
    Vec2 face_center = 0.5f * (v1 + v2);
    float s = b2_dot(c_local - face_center, poly.normals[vi1]);
    if (s > radius) return;
    m.point_count = 1;
    m.type = E_FACE_A;
    m.local_normal = poly.normals[vi1];
    m.local_point = face_center;
    m.points[0].local_point = circle.p;
    m.points[0].id = ContactFeature();
  }
}

// ---- b2_collide_polygon.rs:8-46
inline float find_max_separation(int& edge_index, const Shape& poly1, const Transform& xf1, const Shape& poly2,
                                 const Transform& xf2) {
  int count1 = poly1.count, count2 = poly2.count;
  Transform xf = b2_mul_t_xf_xf(xf2, xf1);
  int best_index = 0;
  float max_separation = -MAX_FLOAT;
  for (int i = 0; i < count1; ++i) {
    Vec2 n = b2_mul_rot(xf.q, poly1.normals[i]);
    Vec2 v1 = b2_mul_xf(xf, poly1.vertices[i]);
    float si = MAX_FLOAT;
    for (int j = 0; j < count2; ++j) {
      float sij = b2_dot(n, poly2.vertices[j] - v1);
      if (sij < si) si = sij;
    }
    if (si > max_separation) { max_separation = si; best_index = i; }
  }
  edge_index = best_index;
  return max_separation;
}

// ---- b2_collide_polygon.rs:48-93
inline void find_incident_edge(ClipVertex c[2], const Shape& poly1, const Transform& xf1, int edge1, const Shape& poly2,
                               const Transform& xf2) {
  int count2 = poly2.count;
  Vec2 normal1 = b2_mul_t_rot(xf2.q, b2_mul_rot(xf1.q, poly1.normals[edge1]));
  int index = 0;
  float min_dot = MAX_FLOAT;
  for (int i = 0; i < count2; ++i) {
    float dot = b2_dot(normal1, poly2.normals[i]);
    if (dot < min_dot) { min_dot = dot; index = i; }
  }
  int i1 = index;
  int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
  c[0].v = b2_mul_xf(xf2, poly2.vertices[i1]);
  c[0].id.index_a = (uint8_t)edge1;
  c[0].id.index_b = (uint8_t)i1;
  c[0].id.type_a = E_FACE;
  c[0].id.type_b = E_VERTEX;
  c[1].v = b2_mul_xf(xf2, poly2.vertices[i2]);
  c[1].id.index_a = (uint8_t)edge1;
  c[1].id.index_b = (uint8_t)i2;
  c[1].id.type_a = E_FACE;
  c[1].id.type_b = E_VERTEX;
}

// ---- b2_collide_polygon.rs:102-225
inline void collide_polygons(Manifold& m, const Shape& poly_a, const Transform& xf_a, const Shape& poly_b,
                             const Transform& xf_b) {
  m.point_count = 0;
  float total_radius = poly_a.radius + poly_b.radius;
  int edge_a = 0;
  float separation_a = find_max_separation(edge_a, poly_a, xf_a, poly_b, xf_b);
  if (separation_a > total_radius) return;
  int edge_b = 0;
  float separation_b = find_max_separation(edge_b, poly_b, xf_b, poly_a, xf_a);
  if (separation_b > total_radius) return;

  const Shape* poly1;
  const Shape* poly2;
  Transform xf1, xf2;
  int edge1;
  uint8_t flip;
  const float k_tol = 0.1f * LINEAR_SLOP;
  if (separation_b > separation_a + k_tol) {
    poly1 = &poly_b; poly2 = &poly_a; xf1 = xf_b; xf2 = xf_a; edge1 = edge_b;
    m.type = E_FACE_B;
    flip = 1;
  } else {
    poly1 = &poly_a; poly2 = &poly_b; xf1 = xf_a; xf2 = xf_b; edge1 = edge_a;
    m.type = E_FACE_A;
    flip = 0;
  }
  ClipVertex incident_edge[2];
  find_incident_edge(incident_edge, *poly1, xf1, edge1, *poly2, xf2);
  int count1 = poly1->count;
  int iv1 = edge1;
  int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  Vec2 v11 = poly1->vertices[iv1], v12 = poly1->vertices[iv2];
  Vec2 local_tangent = v12 - v11;
  local_tangent.normalize();
  Vec2 local_normal = b2_cross_vs(local_tangent, 1.0f);
  Vec2 plane_point = 0.5f * (v11 + v12);
  Vec2 tangent = b2_mul_rot(xf1.q, local_tangent);
  Vec2 normal = b2_cross_vs(tangent, 1.0f);
  v11 = b2_mul_xf(xf1, v11);
  v12 = b2_mul_xf(xf1, v12);
  float front_offset = b2_dot(normal, v11);
  float side_offset1 = -b2_dot(tangent, v11) + total_radius;
  float side_offset2 = b2_dot(tangent, v12) + total_radius;
  ClipVertex clip_points1[2], clip_points2[2];
  int np = clip_segment_to_line(clip_points1, incident_edge, -tangent, side_offset1, iv1);
  if (np < 2) return;
  np = clip_segment_to_line(clip_points2, clip_points1, tangent, side_offset2, iv2);
  if (np < 2) return;
  m.local_normal = local_normal;
  m.local_point = plane_point;
  int point_count = 0;
  for (int i = 0; i < MAX_MANIFOLD_POINTS; ++i) {
    float separation = b2_dot(normal, clip_points2[i].v) - front_offset;
    if (separation <= total_radius) {
      ManifoldPoint& cp = m.points[point_count];
      cp.local_point = b2_mul_t_xf(xf2, clip_points2[i].v);
      cp.id = clip_points2[i].id;
      if (flip) {
        ContactFeature cf = cp.id;
        cp.id.index_a = cf.index_b;
        cp.id.index_b = cf.index_a;
        cp.id.type_a = cf.type_b;
        cp.id.type_b = cf.type_a;
      }
      ++point_count;
    }
  }
  m.point_count = point_count;
}

// ---- b2_collide_edge.rs:11-123
inline void collide_edge_and_circle(Manifold& m, const Shape& edge, const Transform& xf_a, const Shape& circle,
                                    const Transform& xf_b) {
  m.point_count = 0;
  Vec2 q = b2_mul_t_xf(xf_a, b2_mul_xf(xf_b, circle.p));
  Vec2 a = edge.v1, b = edge.v2;
  Vec2 e = b - a;
  Vec2 n(e.y, -e.x);
  float offset = b2_dot(n, q - a);
  bool one_sided = edge.one_sided;
  if (one_sided && offset < 0.0f) return;
  float u = b2_dot(e, b - q);
  float v = b2_dot(e, q - a);
  float radius = edge.radius + circle.radius;
  ContactFeature cf;
  cf.index_b = 0;
  cf.type_b = E_VERTEX;
  if (v <= 0.0f) {
    Vec2 p = a;
    Vec2 d = q - p;
    float dd = b2_dot(d, d);
    if (dd > radius * radius) return;
    if (edge.one_sided) {
      Vec2 a1 = edge.v0, b1 = a;
      Vec2 e1 = b1 - a1;
      float u1 = b2_dot(e1, b1 - q);
      if (u1 > 0.0f) return;
    }
    cf.index_a = 0;
    cf.type_a = E_VERTEX;
    m.point_count = 1;
    m.type = E_CIRCLES;
    m.local_normal.set_zero();
    m.local_point = p;
    m.points[0].id = cf;
    m.points[0].local_point = circle.p;
    return;
  }
  if (u <= 0.0f) {
    Vec2 p = b;
    Vec2 d = q - p;
    float dd = b2_dot(d, d);
    if (dd > radius * radius) return;
    if (edge.one_sided) {
      Vec2 b2 = edge.v3, a2 = b;
      Vec2 e2 = b2 - a2;
      float v2 = b2_dot(e2, q - a2);
      if (v2 > 0.0f) return;
    }
    cf.index_a = 1;
    cf.type_a = E_VERTEX;
    m.point_count = 1;
    m.type = E_CIRCLES;
    m.local_normal.set_zero();
    m.local_point = p;
    m.points[0].id = cf;
    m.points[0].local_point = circle.p;
    return;
  }
  float den = b2_dot(e, e);
  Vec2 p = (1.0f / den) * (u * a + v * b);
  Vec2 d = q - p;
  float dd = b2_dot(d, d);
  if (dd > radius * radius) return;
  if (offset < 0.0f) n.set(-n.x, -n.y);
  n.normalize();
  cf.index_a = 0;
  cf.type_a = E_FACE;
  m.point_count = 1;
  m.type = E_FACE_A;
  m.local_normal = n;
  m.local_point = a;
  m.points[0].id = cf;
  m.points[0].local_point = circle.p;
}

// ---- b2_collide_edge.rs:125-228
enum EPAxisType { EP_UNKNOWN, EP_EDGE_A, EP_EDGE_B };
struct EPAxis {
  Vec2 normal;
  int type = EP_UNKNOWN;
  int index = 0;
  float separation = 0.0f;
};
struct TempPolygon {
  Vec2 vertices[MAX_POLYGON_VERTICES];
  Vec2 normals[MAX_POLYGON_VERTICES];
  int count = 0;
};
struct ReferenceFace {
  int i1 = 0, i2 = 0;
  Vec2 v1, v2, normal, side_normal1;
  float side_offset1 = 0.0f;
  Vec2 side_normal2;
  float side_offset2 = 0.0f;
};
inline EPAxis compute_edge_separation(const TempPolygon& pb, Vec2 v1, Vec2 normal1) {
  EPAxis axis;
  axis.type = EP_EDGE_A;
  axis.index = -1;
  axis.separation = -MAX_FLOAT;
  axis.normal.set_zero();
  Vec2 axes[2] = {normal1, -normal1};
  for (int j = 0; j < 2; ++j) {
    float sj = MAX_FLOAT;
    for (int i = 0; i < pb.count; ++i) {
      float si = b2_dot(axes[j], pb.vertices[i] - v1);
      if (si < sj) sj = si;
    }
    if (sj > axis.separation) { axis.index = j; axis.separation = sj; axis.normal = axes[j]; }
  }
  return axis;
}
inline EPAxis compute_polygon_separation(const TempPolygon& pb, Vec2 v1, Vec2 v2) {
  EPAxis axis;
  axis.type = EP_UNKNOWN;
  axis.index = -1;
  axis.separation = -MAX_FLOAT;
  axis.normal.set_zero();
  for (int i = 0; i < pb.count; ++i) {
    Vec2 n = -pb.normals[i];
    float s1 = b2_dot(n, pb.vertices[i] - v1);
    float s2 = b2_dot(n, pb.vertices[i] - v2);
    float s = b2_min(s1, s2);
    if (s > axis.separation) { axis.type = EP_EDGE_B; axis.index = i; axis.separation = s; axis.normal = n; }
  }
  return axis;
}

// ---- b2_collide_edge.rs:230-475
inline void collide_edge_and_polygon(Manifold& m, const Shape& edge, const Transform& xf_a, const Shape& poly_b,
                                     const Transform& xf_b) {
  m.point_count = 0;
  Transform xf = b2_mul_t_xf_xf(xf_a, xf_b);
  Vec2 centroid_b = b2_mul_xf(xf, poly_b.centroid);
  Vec2 v1 = edge.v1, v2 = edge.v2;
  Vec2 edge1 = v2 - v1;
  edge1.normalize();
  Vec2 normal1(edge1.y, -edge1.x);
  float offset1 = b2_dot(normal1, centroid_b - v1);
  bool one_sided = edge.one_sided;
  if (one_sided && offset1 < 0.0f) return;
  TempPolygon tp;
  tp.count = poly_b.count;
  for (int i = 0; i < poly_b.count; ++i) {
    tp.vertices[i] = b2_mul_xf(xf, poly_b.vertices[i]);
    tp.normals[i] = b2_mul_rot(xf.q, poly_b.normals[i]);
  }
  float radius = poly_b.radius + edge.radius;
  EPAxis edge_axis = compute_edge_separation(tp, v1, normal1);
  if (edge_axis.separation > radius) return;
  EPAxis polygon_axis = compute_polygon_separation(tp, v1, v2);
  if (polygon_axis.separation > radius) return;
  const float k_relative_tol = 0.98f, k_absolute_tol = 0.001f;
  EPAxis primary_axis;
  if (polygon_axis.separation - radius > k_relative_tol * (edge_axis.separation - radius) + k_absolute_tol)
    primary_axis = polygon_axis;
  else
    primary_axis = edge_axis;
  if (one_sided) {
    Vec2 edge0 = v1 - edge.v0;
    edge0.normalize();
    Vec2 normal0(edge0.y, -edge0.x);
    bool convex1 = b2_cross(edge0, edge1) >= 0.0f;
    Vec2 edge2 = edge.v3 - v2;
    edge2.normalize();
    Vec2 normal2(edge2.y, -edge2.x);
    bool convex2 = b2_cross(edge1, edge2) >= 0.0f;
    const float sin_tol = 0.1f;
    bool side1 = b2_dot(primary_axis.normal, edge1) <= 0.0f;
    if (side1) {
      if (convex1) {
        if (b2_cross(primary_axis.normal, normal0) > sin_tol) return;
      } else {
        primary_axis = edge_axis;
      }
    } else {
      if (convex2) {
        if (b2_cross(normal2, primary_axis.normal) > sin_tol) return;
      } else {
        primary_axis = edge_axis;
      }
    }
  }
  ClipVertex clip_points[2];
  ReferenceFace rf;
  if (primary_axis.type == EP_EDGE_A) {
    m.type = E_FACE_A;
    int best_index = 0;
    float best_value = b2_dot(primary_axis.normal, tp.normals[0]);
    for (int i = 1; i < tp.count; ++i) {
      float value = b2_dot(primary_axis.normal, tp.normals[i]);
      if (value < best_value) { best_value = value; best_index = i; }
    }
    int i1 = best_index;
    int i2 = i1 + 1 < tp.count ? i1 + 1 : 0;
    clip_points[0].v = tp.vertices[i1];
    clip_points[0].id.index_a = 0;
    clip_points[0].id.index_b = (uint8_t)i1;
    clip_points[0].id.type_a = E_FACE;
    clip_points[0].id.type_b = E_VERTEX;
    clip_points[1].v = tp.vertices[i2];
    clip_points[1].id.index_a = 0;
    clip_points[1].id.index_b = (uint8_t)i2;
    clip_points[1].id.type_a = E_FACE;
    clip_points[1].id.type_b = E_VERTEX;
    rf.i1 = 0;
    rf.i2 = 1;
    rf.v1 = v1;
    rf.v2 = v2;
    rf.normal = primary_axis.normal;
    rf.side_normal1 = -edge1;
    rf.side_normal2 = edge1;
  } else {
    m.type = E_FACE_B;
    clip_points[0].v = v2;
    clip_points[0].id.index_a = 1;
    clip_points[0].id.index_b = (uint8_t)primary_axis.index;
    clip_points[0].id.type_a = E_VERTEX;
    clip_points[0].id.type_b = E_FACE;
    clip_points[1].v = v1;
    clip_points[1].id.index_a = 0;
    clip_points[1].id.index_b = (uint8_t)primary_axis.index;
    clip_points[1].id.type_a = E_VERTEX;
    clip_points[1].id.type_b = E_FACE;
    rf.i1 = primary_axis.index;
    rf.i2 = rf.i1 + 1 < tp.count ? rf.i1 + 1 : 0;
    rf.v1 = tp.vertices[rf.i1];
    rf.v2 = tp.vertices[rf.i2];
    rf.normal = tp.normals[rf.i1];
    rf.side_normal1.set(rf.normal.y, -rf.normal.x);
    rf.side_normal2 = -rf.side_normal1;
  }
  rf.side_offset1 = b2_dot(rf.side_normal1, rf.v1);
  rf.side_offset2 = b2_dot(rf.side_normal2, rf.v2);
  ClipVertex clip_points1[2], clip_points2[2];
  int np = clip_segment_to_line(clip_points1, clip_points, rf.side_normal1, rf.side_offset1, rf.i1);
  if (np < MAX_MANIFOLD_POINTS) return;
  np = clip_segment_to_line(clip_points2, clip_points1, rf.side_normal2, rf.side_offset2, rf.i2);
  if (np < MAX_MANIFOLD_POINTS) return;
  if (primary_axis.type == EP_EDGE_A) {
    m.local_normal = rf.normal;
    m.local_point = rf.v1;
  } else {
    m.local_normal = poly_b.normals[rf.i1];
    m.local_point = poly_b.vertices[rf.i1];
  }
  int point_count = 0;
  for (int i = 0; i < MAX_MANIFOLD_POINTS; ++i) {
    float separation = b2_dot(rf.normal, clip_points2[i].v - rf.v1);
    if (separation <= radius) {
      ManifoldPoint& cp = m.points[point_count];
      if (primary_axis.type == EP_EDGE_A) {
        cp.local_point = b2_mul_t_xf(xf, clip_points2[i].v);
        cp.id = clip_points2[i].id;
      } else {
        cp.local_point = clip_points2[i].v;
        cp.id.type_a = clip_points2[i].id.type_b;
        cp.id.type_b = clip_points2[i].id.type_a;
        cp.id.index_a = clip_points2[i].id.index_b;
        cp.id.index_b = clip_points2[i].id.index_a;
      }
      ++point_count;
    }
  }
  m.point_count = point_count;
}

}  // namespace b2o
