// b2g_distance.h — GJK overlap test for sensor contacts.
//
// B2contact::update decides `touching` of a sensor contact with b2_test_overlap on the two shapes
// (b2_contact.rs(private):149-163 -> src/private/collision/b2_collision.rs:216-243): b2_distance_fn with
// use_radii and an empty simplex cache, overlap iff distance < 10 * epsilon.  This restates
// src/private/collision/b2_distance.rs (B2simplex :67-258, solve2 :282-310, solve3 :317-409,
// b2_distance_fn :411-540) and get_support (src/b2_distance.rs:143-155) on the flat shape records,
// operation for operation.  Sensor contacts are rare: the device function is not inlined, so the
// manifold path of CollideK keeps its register budget.
#pragma once
#include "b2g_narrow.h"

namespace b2g {

struct GjkProxy {
  V2 v[8];
  int count;
  float radius;
};
B2G_HD void gjk_proxy_set(GjkProxy& p, ShapeP s) {  // set_shape :11-44 (chain children arrive as edge records)
  if (s->type == B2GPU_SHAPE_CIRCLE) {
    p.v[0] = v2(s->cx, s->cy);
    p.count = 1;
  } else if (s->type == B2GPU_SHAPE_POLYGON) {
    for (int i = 0; i < s->count; ++i) p.v[i] = sh_vert(s, i);
    p.count = s->count;
  } else {
    p.v[0] = sh_vert(s, 1);  // m_vertex1
    p.v[1] = sh_vert(s, 2);  // m_vertex2
    p.count = 2;
  }
  p.radius = s->radius;
}
B2G_HD int gjk_support(const GjkProxy& p, V2 d) {
  int best_index = 0;
  float best_value = dot(p.v[0], d);
  for (int i = 1; i < p.count; ++i) {
    const float value = dot(p.v[i], d);
    if (value > best_value) {
      best_index = i;
      best_value = value;
    }
  }
  return best_index;
}

struct GjkVertex {
  V2 w_a, w_b, w;
  float a;
  int index_a, index_b;
};
struct GjkSimplex {
  GjkVertex v[3];
  int count;
};

B2G_HD void gjk_solve2(GjkSimplex& s) {
  const V2 w1 = s.v[0].w, w2 = s.v[1].w;
  const V2 e12 = w2 - w1;
  const float d12_2 = -dot(w1, e12);
  if (d12_2 <= 0.0f) {
    s.v[0].a = 1.0f;
    s.count = 1;
    return;
  }
  const float d12_1 = dot(w2, e12);
  if (d12_1 <= 0.0f) {
    s.v[1].a = 1.0f;
    s.count = 1;
    s.v[0] = s.v[1];
    return;
  }
  const float inv_d12 = 1.0f / (d12_1 + d12_2);
  s.v[0].a = d12_1 * inv_d12;
  s.v[1].a = d12_2 * inv_d12;
  s.count = 2;
}

B2G_HD void gjk_solve3(GjkSimplex& s) {
  const V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
  const V2 e12 = w2 - w1;
  const float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
  const float d12_1 = w2e12, d12_2 = -w1e12;
  const V2 e13 = w3 - w1;
  const float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
  const float d13_1 = w3e13, d13_2 = -w1e13;
  const V2 e23 = w3 - w2;
  const float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
  const float d23_1 = w3e23, d23_2 = -w2e23;
  const float n123 = cross(e12, e13);
  const float d123_1 = n123 * cross(w2, w3);
  const float d123_2 = n123 * cross(w3, w1);
  const float d123_3 = n123 * cross(w1, w2);
  if (d12_2 <= 0.0f && d13_2 <= 0.0f) {  // w1 region
    s.v[0].a = 1.0f;
    s.count = 1;
    return;
  }
  if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {  // e12
    const float inv_d12 = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv_d12;
    s.v[1].a = d12_2 * inv_d12;
    s.count = 2;
    return;
  }
  if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {  // e13
    const float inv_d13 = 1.0f / (d13_1 + d13_2);
    s.v[0].a = d13_1 * inv_d13;
    s.v[2].a = d13_2 * inv_d13;
    s.count = 2;
    s.v[1] = s.v[2];
    return;
  }
  if (d12_1 <= 0.0f && d23_2 <= 0.0f) {  // w2 region
    s.v[1].a = 1.0f;
    s.count = 1;
    s.v[0] = s.v[1];
    return;
  }
  if (d13_1 <= 0.0f && d23_1 <= 0.0f) {  // w3 region
    s.v[2].a = 1.0f;
    s.count = 1;
    s.v[0] = s.v[2];
    return;
  }
  if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {  // e23
    const float inv_d23 = 1.0f / (d23_1 + d23_2);
    s.v[1].a = d23_1 * inv_d23;
    s.v[2].a = d23_2 * inv_d23;
    s.count = 2;
    s.v[0] = s.v[2];
    return;
  }
  const float inv_d123 = 1.0f / (d123_1 + d123_2 + d123_3);  // inside the triangle
  s.v[0].a = d123_1 * inv_d123;
  s.v[1].a = d123_2 * inv_d123;
  s.v[2].a = d123_3 * inv_d123;
  s.count = 3;
}

// b2_test_overlap (shapes): distance with radii from an empty cache, < 10 * epsilon.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
bool test_overlap_shapes(ShapeP shape_a, ShapeP shape_b, const Xf xf_a, const Xf xf_b) {
  GjkProxy pa, pb;
  gjk_proxy_set(pa, shape_a);
  gjk_proxy_set(pb, shape_b);
  GjkSimplex simplex;
  {  // read_cache with count == 0 (:112-124)
    GjkVertex& v = simplex.v[0];
    v.index_a = 0;
    v.index_b = 0;
    v.w_a = xf_mul(xf_a, pa.v[0]);
    v.w_b = xf_mul(xf_b, pb.v[0]);
    v.w = v.w_b - v.w_a;
    v.a = 1.0f;
    simplex.count = 1;
  }
  int save_a[3] = {0, 0, 0}, save_b[3] = {0, 0, 0};
  int iter = 0;
  while (iter < 20) {
    const int save_count = simplex.count;
    for (int i = 0; i < save_count; ++i) {
      save_a[i] = simplex.v[i].index_a;
      save_b[i] = simplex.v[i].index_b;
    }
    if (simplex.count == 2) gjk_solve2(simplex);
    else if (simplex.count == 3) gjk_solve3(simplex);
    if (simplex.count == 3) break;
    V2 d;  // get_search_direction :138-162
    if (simplex.count == 1) {
      d = -simplex.v[0].w;
    } else {
      const V2 e12 = simplex.v[1].w - simplex.v[0].w;
      const float sgn = cross(e12, -simplex.v[0].w);
      d = sgn > 0.0f ? cross_sv(1.0f, e12) : cross_vs(e12, 1.0f);
    }
    if (dot(d, d) < B2G_EPSILON * B2G_EPSILON) break;
    GjkVertex& vertex = simplex.v[simplex.count];
    vertex.index_a = gjk_support(pa, rot_mul_t(xf_a.q, -d));
    vertex.w_a = xf_mul(xf_a, pa.v[vertex.index_a]);
    vertex.index_b = gjk_support(pb, rot_mul_t(xf_b.q, d));
    vertex.w_b = xf_mul(xf_b, pb.v[vertex.index_b]);
    vertex.w = vertex.w_b - vertex.w_a;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < save_count; ++i)
      if (vertex.index_a == save_a[i] && vertex.index_b == save_b[i]) {
        duplicate = true;
        break;
      }
    if (duplicate) break;
    ++simplex.count;
  }
  V2 point_a, point_b;  // get_witness_points :189-228
  if (simplex.count == 1) {
    point_a = simplex.v[0].w_a;
    point_b = simplex.v[0].w_b;
  } else if (simplex.count == 2) {
    point_a = simplex.v[0].a * simplex.v[0].w_a + simplex.v[1].a * simplex.v[1].w_a;
    point_b = simplex.v[0].a * simplex.v[0].w_b + simplex.v[1].a * simplex.v[1].w_b;
  } else {
    point_a = simplex.v[0].a * simplex.v[0].w_a + simplex.v[1].a * simplex.v[1].w_a + simplex.v[2].a * simplex.v[2].w_a;
    point_b = point_a;
  }
  float distance = length(point_a - point_b);
  const float r_a = pa.radius, r_b = pb.radius;
  if (distance > r_a + r_b && distance > B2G_EPSILON) distance -= r_a + r_b;
  else distance = 0.0f;
  return distance < 10.0f * B2G_EPSILON;
}

}  // namespace b2g
