// TEST INFRASTRUCTURE — CPU oracle (see b2o_math.hpp header). PARITY UNPINNED beyond the
// reference's own tests.
//
// b2o_distance.hpp — restates the GJK distance query of src/private/collision/b2_distance.rs
// (set_shape :11-44, B2simplex :67-258, b2_simplex_solve2 :282-310, b2_simplex_solve3 :317-409,
// b2_distance_fn :411-540), get_support src/b2_distance.rs:143-155, and
// b2_test_overlap (shapes) src/private/collision/b2_collision.rs:216-243, which is what
// B2contact::update calls for sensor contacts (b2_contact.rs(private):149-163).
#pragma once
#include "b2o_collision.hpp"

namespace b2o {

struct DistanceProxy {
  Vec2 vertices[MAX_POLYGON_VERTICES];
  int count = 0;
  float radius = 0.0f;
  int get_support(Vec2 d) const {  // src/b2_distance.rs:143-155
    int best_index = 0;
    float best_value = b2_dot(vertices[0], d);
    for (int i = 1; i < count; ++i) {
      float value = b2_dot(vertices[i], d);
      if (value > best_value) {
        best_index = i;
        best_value = value;
      }
    }
    return best_index;
  }
};

// set_shape :11-44 (the chain case takes child `index`; a chain child edge always has index + 1 < len)
inline void proxy_set_shape(DistanceProxy& p, const Shape& s, int index) {
  switch (s.type) {
    case E_CIRCLE:
      p.vertices[0] = s.p;
      p.count = 1;
      p.radius = s.radius;
      break;
    case E_POLYGON:
      for (int i = 0; i < s.count; ++i) p.vertices[i] = s.vertices[i];
      p.count = s.count;
      p.radius = s.radius;
      break;
    case E_CHAIN:
      assert(index < (int)s.chain.size());
      p.vertices[0] = s.chain[index];
      p.vertices[1] = index + 1 < (int)s.chain.size() ? s.chain[index + 1] : s.chain[0];
      p.count = 2;
      p.radius = s.radius;
      break;
    default:  // E_EDGE
      p.vertices[0] = s.v1;
      p.vertices[1] = s.v2;
      p.count = 2;
      p.radius = s.radius;
      break;
  }
}

struct SimplexCache {
  float metric = 0.0f;
  int count = 0;
  int index_a[3] = {0, 0, 0}, index_b[3] = {0, 0, 0};
};
struct SimplexVertex {
  Vec2 w_a, w_b, w;
  float a = 0.0f;
  int index_a = 0, index_b = 0;
};
struct Simplex {
  SimplexVertex v[3];
  int count = 0;

  float get_metric() const {  // :230-250
    switch (count) {
      case 1: return 0.0f;
      case 2: return (v[0].w - v[1].w).length();
      case 3: return b2_cross(v[1].w - v[0].w, v[2].w - v[0].w);
      default: assert(false); return 0.0f;
    }
  }
  void read_cache(const SimplexCache& cache, const DistanceProxy& pa, const Transform& xa, const DistanceProxy& pb,
                  const Transform& xb) {  // :74-126
    count = cache.count;
    for (int i = 0; i < count; ++i) {
      SimplexVertex& s = v[i];
      s.index_a = cache.index_a[i];
      s.index_b = cache.index_b[i];
      s.w_a = b2_mul_xf(xa, pa.vertices[s.index_a]);
      s.w_b = b2_mul_xf(xb, pb.vertices[s.index_b]);
      s.w = s.w_b - s.w_a;
      s.a = 0.0f;
    }
    if (count > 1) {
      float metric1 = cache.metric, metric2 = get_metric();
      if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < EPSILON) count = 0;
    }
    if (count == 0) {
      SimplexVertex& s = v[0];
      s.index_a = 0;
      s.index_b = 0;
      s.w_a = b2_mul_xf(xa, pa.vertices[0]);
      s.w_b = b2_mul_xf(xb, pb.vertices[0]);
      s.w = s.w_b - s.w_a;
      s.a = 1.0f;
      count = 1;
    }
  }
  void write_cache(SimplexCache& cache) const {  // :128-136
    cache.metric = get_metric();
    cache.count = count;
    for (int i = 0; i < count; ++i) {
      cache.index_a[i] = v[i].index_a;
      cache.index_b[i] = v[i].index_b;
    }
  }
  Vec2 get_search_direction() const {  // :138-162
    if (count == 1) return -v[0].w;
    assert(count == 2);
    Vec2 e12 = v[1].w - v[0].w;
    float sgn = b2_cross(e12, -v[0].w);
    if (sgn > 0.0f) return b2_cross_sv(1.0f, e12);  // origin is left of e12
    return b2_cross_vs(e12, 1.0f);                   // origin is right of e12
  }
  void get_witness_points(Vec2& p_a, Vec2& p_b) const {  // :189-228
    switch (count) {
      case 1:
        p_a = v[0].w_a;
        p_b = v[0].w_b;
        break;
      case 2:
        p_a = v[0].a * v[0].w_a + v[1].a * v[1].w_a;
        p_b = v[0].a * v[0].w_b + v[1].a * v[1].w_b;
        break;
      case 3:
        p_a = v[0].a * v[0].w_a + v[1].a * v[1].w_a + v[2].a * v[2].w_a;
        p_b = p_a;
        break;
      default: assert(false);
    }
  }
  void solve2() {  // :282-310
    Vec2 w1 = v[0].w, w2 = v[1].w;
    Vec2 e12 = w2 - w1;
    float d12_2 = -b2_dot(w1, e12);
    if (d12_2 <= 0.0f) {
      v[0].a = 1.0f;
      count = 1;
      return;
    }
    float d12_1 = b2_dot(w2, e12);
    if (d12_1 <= 0.0f) {
      v[1].a = 1.0f;
      count = 1;
      v[0] = v[1];
      return;
    }
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    v[0].a = d12_1 * inv_d12;
    v[1].a = d12_2 * inv_d12;
    count = 2;
  }
  void solve3() {  // :317-409
    Vec2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
    Vec2 e12 = w2 - w1;
    float w1e12 = b2_dot(w1, e12), w2e12 = b2_dot(w2, e12);
    float d12_1 = w2e12, d12_2 = -w1e12;
    Vec2 e13 = w3 - w1;
    float w1e13 = b2_dot(w1, e13), w3e13 = b2_dot(w3, e13);
    float d13_1 = w3e13, d13_2 = -w1e13;
    Vec2 e23 = w3 - w2;
    float w2e23 = b2_dot(w2, e23), w3e23 = b2_dot(w3, e23);
    float d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = b2_cross(e12, e13);
    float d123_1 = n123 * b2_cross(w2, w3);
    float d123_2 = n123 * b2_cross(w3, w1);
    float d123_3 = n123 * b2_cross(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) {  // w1 region
      v[0].a = 1.0f;
      count = 1;
      return;
    }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {  // e12
      float inv_d12 = 1.0f / (d12_1 + d12_2);
      v[0].a = d12_1 * inv_d12;
      v[1].a = d12_2 * inv_d12;
      count = 2;
      return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {  // e13
      float inv_d13 = 1.0f / (d13_1 + d13_2);
      v[0].a = d13_1 * inv_d13;
      v[2].a = d13_2 * inv_d13;
      count = 2;
      v[1] = v[2];
      return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) {  // w2 region
      v[1].a = 1.0f;
      count = 1;
      v[0] = v[1];
      return;
    }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) {  // w3 region
      v[2].a = 1.0f;
      count = 1;
      v[0] = v[2];
      return;
    }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {  // e23
      float inv_d23 = 1.0f / (d23_1 + d23_2);
      v[1].a = d23_1 * inv_d23;
      v[2].a = d23_2 * inv_d23;
      count = 2;
      v[0] = v[2];
      return;
    }
    float inv_d123 = 1.0f / (d123_1 + d123_2 + d123_3);  // inside the triangle
    v[0].a = d123_1 * inv_d123;
    v[1].a = d123_2 * inv_d123;
    v[2].a = d123_3 * inv_d123;
    count = 3;
  }
};

struct DistanceOutput {
  Vec2 point_a, point_b;
  float distance = 0.0f;
  int iterations = 0;
};

// b2_distance_fn :411-540
inline void b2_distance(DistanceOutput& output, SimplexCache& cache, const DistanceProxy& proxy_a, const Transform& transform_a,
                        const DistanceProxy& proxy_b, const Transform& transform_b, bool use_radii) {
  Simplex simplex;
  simplex.read_cache(cache, proxy_a, transform_a, proxy_b, transform_b);
  const int k_max_iters = 20;
  int save_a[3] = {0, 0, 0}, save_b[3] = {0, 0, 0};
  int save_count;
  int iter = 0;
  while (iter < k_max_iters) {
    save_count = simplex.count;
    for (int i = 0; i < save_count; ++i) {
      save_a[i] = simplex.v[i].index_a;
      save_b[i] = simplex.v[i].index_b;
    }
    switch (simplex.count) {
      case 1: break;
      case 2: simplex.solve2(); break;
      case 3: simplex.solve3(); break;
      default: assert(false);
    }
    if (simplex.count == 3) break;
    Vec2 d = simplex.get_search_direction();
    if (d.length_squared() < EPSILON * EPSILON) break;
    SimplexVertex& vertex = simplex.v[simplex.count];
    vertex.index_a = proxy_a.get_support(b2_mul_t_rot(transform_a.q, -d));
    vertex.w_a = b2_mul_xf(transform_a, proxy_a.vertices[vertex.index_a]);
    vertex.index_b = proxy_b.get_support(b2_mul_t_rot(transform_b.q, d));
    vertex.w_b = b2_mul_xf(transform_b, proxy_b.vertices[vertex.index_b]);
    vertex.w = vertex.w_b - vertex.w_a;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < save_count; ++i) {
      if (vertex.index_a == save_a[i] && vertex.index_b == save_b[i]) {
        duplicate = true;
        break;
      }
    }
    if (duplicate) break;
    ++simplex.count;
  }
  simplex.get_witness_points(output.point_a, output.point_b);
  output.distance = (output.point_a - output.point_b).length();
  output.iterations = iter;
  simplex.write_cache(cache);
  if (use_radii) {
    float r_a = proxy_a.radius, r_b = proxy_b.radius;
    if (output.distance > r_a + r_b && output.distance > EPSILON) {
      output.distance -= r_a + r_b;
      Vec2 normal = output.point_b - output.point_a;
      normal.normalize();
      output.point_a += r_a * normal;
      output.point_b -= r_b * normal;
    } else {
      Vec2 p = 0.5f * (output.point_a + output.point_b);
      output.point_a = p;
      output.point_b = p;
      output.distance = 0.0f;
    }
  }
}

// b2_test_overlap (shapes) b2_collision.rs(private):216-243
inline bool b2_test_overlap_shapes(const Shape& shape_a, int index_a, const Shape& shape_b, int index_b, const Transform& xf_a,
                                   const Transform& xf_b) {
  DistanceProxy pa, pb;
  proxy_set_shape(pa, shape_a, index_a);
  proxy_set_shape(pb, shape_b, index_b);
  SimplexCache cache;
  DistanceOutput output;
  b2_distance(output, cache, pa, xf_a, pb, xf_b, true);
  return output.distance < 10.0f * EPSILON;
}

}  // namespace b2o
