// b2g_query.h — world queries on the device: many rays / boxes at once, one thread each.
//
// Reference: B2world::ray_cast (src/private/dynamics/b2_world.rs:1015-1049) -> B2dynamicTree::ray_cast
// (src/b2_dynamic_tree.rs:269-347) -> B2fixture::ray_cast (src/b2_fixture.rs:228) -> the shape ray casts
// (b2_circle_shape.rs(private):26-62, b2_edge_shape.rs(private):39-102, b2_polygon_shape.rs(private):226-290,
// b2_chain_shape.rs(private):86-108); B2world::query_aabb (:969-980) -> B2dynamicTree::query (:239-267).
// The ray cast implements the callback `|fixture, point, normal, fraction| fraction` (closest hit: the ray is
// clipped to every hit, so the result does not depend on the traversal order except between exactly equal
// fractions); query_aabb implements a callback that always continues.  In the exact modes the walk is the
// reference's walk over the replica tree (same report order); in large-world mode 1 it is the LBVH.
#pragma once
#include "b2g_large.h"

namespace b2g {

struct RayIn { V2 p1, p2; float max_fraction; };

B2G_HD bool ray_cast_circle(ShapeP s, const Xf& xf, const RayIn& in, float& fraction, V2& normal) {
  const V2 position = xf.p + rot_mul(xf.q, sh_center(s));
  const V2 sv = in.p1 - position;
  const float b = dot(sv, sv) - s->radius * s->radius;
  const V2 r = in.p2 - in.p1;
  const float c = dot(sv, r);
  const float rr = dot(r, r);
  const float sigma = c * c - rr * b;
  if (sigma < 0.0f || rr < B2G_EPSILON) return false;
  float a = -(c + sqrtf(sigma));
  if (0.0f <= a && a <= in.max_fraction * rr) {
    a /= rr;
    fraction = a;
    normal = sv + a * r;
    normalize(normal);
    return true;
  }
  return false;
}
B2G_HD bool ray_cast_edge(V2 v1, V2 v2_, bool one_sided, const Xf& xf, const RayIn& in, float& fraction, V2& normal_out) {
  const V2 p1 = rot_mul_t(xf.q, in.p1 - xf.p);
  const V2 p2 = rot_mul_t(xf.q, in.p2 - xf.p);
  const V2 d = p2 - p1;
  const V2 e = v2_ - v1;
  V2 normal = v2(e.y, -e.x);
  normalize(normal);
  const float numerator = dot(normal, v1 - p1);
  if (one_sided && numerator > 0.0f) return false;
  const float denominator = dot(normal, d);
  if (denominator == 0.0f) return false;
  const float t = numerator / denominator;
  if (t < 0.0f || in.max_fraction < t) return false;
  const V2 q = p1 + t * d;
  const V2 r = v2_ - v1;
  const float rr = dot(r, r);
  if (rr == 0.0f) return false;
  const float s = dot(q - v1, r) / rr;
  if (s < 0.0f || 1.0f < s) return false;
  fraction = t;
  if (numerator > 0.0f) normal_out = -rot_mul(xf.q, normal);
  else normal_out = rot_mul(xf.q, normal);
  return true;
}
B2G_HD bool ray_cast_polygon(ShapeP s, const Xf& xf, const RayIn& in, float& fraction, V2& normal) {
  const V2 p1 = rot_mul_t(xf.q, in.p1 - xf.p);
  const V2 p2 = rot_mul_t(xf.q, in.p2 - xf.p);
  const V2 d = p2 - p1;
  float lower = 0.032f, upper = in.max_fraction;  // sic: the Rust port starts at 0.032 (b2_polygon_shape.rs(private):241)
  int index = -1;
  for (int i = 0; i < s->count; ++i) {
    const float numerator = dot(sh_norm(s, i), sh_vert(s, i) - p1);
    const float denominator = dot(sh_norm(s, i), d);
    if (denominator == 0.0f) {
      if (numerator < 0.0f) return false;
    } else {
      if (denominator < 0.0f && numerator < lower * denominator) {
        lower = numerator / denominator;
        index = i;
      } else if (denominator > 0.0f && numerator < upper * denominator) {
        upper = numerator / denominator;
      }
    }
    if (upper < lower) return false;
  }
  if (index >= 0) {
    fraction = lower;
    normal = rot_mul(xf.q, sh_norm(s, index));
    return true;
  }
  return false;
}
// B2fixture::ray_cast of one child; chains cast against a default (two-sided) edge of their vertices i, i+1
B2G_HD bool ray_cast_child(const Batch& B, const WIdx& x, int fixture, int child, const RayIn& in, float& fraction, V2& normal) {
  const b2gpu_fixture_rec& f = B.fixtures[fixture];
  ShapeP s = &B.shapes[f.shape_first + child];
  const Xf xf = load_xf(B, x, f.body);
  if (s->type == B2GPU_SHAPE_CIRCLE) return ray_cast_circle(s, xf, in, fraction, normal);
  if (s->type == B2GPU_SHAPE_POLYGON) return ray_cast_polygon(s, xf, in, fraction, normal);
  return ray_cast_edge(sh_vert(s, 1), sh_vert(s, 2), f.shape_type == B2GPU_SHAPE_CHAIN ? false : s->one_sided != 0, xf, in, fraction, normal);
}

struct RayCastK {  // flat over (world, ray): tid = world * rays + r (any LB)
  Batch B;
  Large L;
  const float* rays;    // [n_worlds][rays][4]
  b2gpu_ray_hit* out;   // [n_worlds][rays]
  int n_rays, use_lbvh, n_leaves;
  B2G_HD void operator()(int tid) const {
    const int w = tid / n_rays;
    if (w >= B.n_worlds) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    const float* rp = rays + (size_t)tid * 4;
    const V2 p1 = v2(rp[0], rp[1]), p2 = v2(rp[2], rp[3]);
    b2gpu_ray_hit hit;
    hit.fixture = -1; hit.child_index = 0; hit.fraction = 0.0f;
    hit.point_x = hit.point_y = hit.normal_x = hit.normal_y = 0.0f;
    hit.reserved = 0;
    V2 r = p2 - p1;
    normalize(r);
    const V2 v = cross_sv(1.0f, r);
    const V2 abs_v = v2(fabsf(v.x), fabsf(v.y));
    float max_fraction = 1.0f;
    Box seg;
    {
      const V2 t = p1 + max_fraction * (p2 - p1);
      seg.lo = vmin(p1, t);
      seg.hi = vmax(p1, t);
    }
    int stack[LW_STACK];
    int sp_ = 0;
    // node references: replica tree ids (>= 0, -1 = null); LBVH: >= 0 internal node, < 0 leaf at sorted position ~c
    stack[sp_++] = use_lbvh ? (n_leaves > 1 ? 0 : ~0) : ws[WS_TREE_ROOT];
    if (use_lbvh && n_leaves < 1) sp_ = 0;
    while (sp_ > 0) {
      const int id = stack[--sp_];
      int leaf_node = -1;  // tree node id of a leaf to test
      Box nb;
      if (use_lbvh) {
        if (id < 0) { leaf_node = L.lb_leaf[~id]; nb = load_box(B.n_aabb, x.at(B.NN, leaf_node)); }
        else nb = load_box(L.lb_box, id);
      } else {
        if (id == -1) continue;
        nb = load_box(B.n_aabb, x.at(B.NN, id));
      }
      if (!box_overlap(nb, seg)) continue;
      const V2 c = box_center(nb);
      const V2 h = 0.5f * (nb.hi - nb.lo);
      const float separation = fabsf(dot(v, p1 - c)) - dot(abs_v, h);
      if (separation > 0.0f) continue;
      if (!use_lbvh) {
        const int4 l = B.n_link[x.at(B.NN, id)];
        if (l.y == -1) leaf_node = id;
        else {
          if (sp_ + 2 > LW_STACK) { ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
          stack[sp_++] = l.y;
          stack[sp_++] = l.z;
          continue;
        }
      } else if (id >= 0) {
        const int2 ch = L.lb_child[id];
        if (sp_ + 2 > LW_STACK) { ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
        stack[sp_++] = ch.y;
        stack[sp_++] = ch.x;
        continue;
      }
      // leaf: the ray-cast callback of b2_world.rs(private):1031-1048
      const int4 ps = B.proxy_s[B.node_proxy[leaf_node]];
      RayIn in;
      in.p1 = p1; in.p2 = p2; in.max_fraction = max_fraction;
      float fraction;
      V2 normal;
      float value = in.max_fraction;
      if (ray_cast_child(B, x, ps.x, ps.y, in, fraction, normal)) {
        const V2 point = (1.0f - fraction) * p1 + fraction * p2;
        hit.fixture = ps.x; hit.child_index = ps.y; hit.fraction = fraction;
        hit.point_x = point.x; hit.point_y = point.y; hit.normal_x = normal.x; hit.normal_y = normal.y;
        value = fraction;
      }
      if (value == 0.0f) break;  // the client has terminated the ray cast
      if (value > 0.0f) {
        max_fraction = value;
        const V2 t = p1 + max_fraction * (p2 - p1);
        seg.lo = vmin(p1, t);
        seg.hi = vmax(p1, t);
      }
    }
    out[tid] = hit;
  }
};

struct QueryAabbK {  // flat over boxes: all of world 0 (per_world == 0), or per_world boxes of every world of a batch
  Batch B;
  Large L;
  const float* boxes;  // [n][4]
  int* counts;         // [n]
  int* hits;           // [n][max_hits][2]
  int n, max_hits, use_lbvh, n_leaves;
  int per_world;       // 0: every box queries world 0; k > 0: box i queries world i / k
  B2G_HD void operator()(int i) const {
    if (i >= n) return;
    WIdx x = widx(B, per_world > 0 ? i / per_world : 0);
    Ws ws = ws_of(B, x);
    Box q;
    q.lo = v2(boxes[4 * i], boxes[4 * i + 1]);
    q.hi = v2(boxes[4 * i + 2], boxes[4 * i + 3]);
    int count = 0;
    int stack[LW_STACK];
    int sp_ = 0;
    stack[sp_++] = use_lbvh ? (n_leaves > 1 ? 0 : ~0) : ws[WS_TREE_ROOT];
    if (use_lbvh && n_leaves < 1) sp_ = 0;
    while (sp_ > 0) {
      const int id = stack[--sp_];
      int leaf_node = -1;
      if (use_lbvh) {
        if (id < 0) {
          leaf_node = L.lb_leaf[~id];
          if (!box_overlap(load_box(B.n_aabb, x.at(B.NN, leaf_node)), q)) continue;
        } else {
          if (!box_overlap(load_box(L.lb_box, id), q)) continue;
          const int2 ch = L.lb_child[id];
          if (sp_ + 2 > LW_STACK) { ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
          stack[sp_++] = ch.y;
          stack[sp_++] = ch.x;
          continue;
        }
      } else {
        if (id == -1) continue;
        if (!box_overlap(load_box(B.n_aabb, x.at(B.NN, id)), q)) continue;
        const int4 l = B.n_link[x.at(B.NN, id)];
        if (l.y != -1) {
          if (sp_ + 2 > LW_STACK) { ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
          stack[sp_++] = l.y;
          stack[sp_++] = l.z;
          continue;
        }
        leaf_node = id;
      }
      const int4 ps = B.proxy_s[B.node_proxy[leaf_node]];
      if (count < max_hits) {
        hits[((size_t)i * max_hits + count) * 2] = ps.x;
        hits[((size_t)i * max_hits + count) * 2 + 1] = ps.y;
      }
      ++count;
    }
    counts[i] = count;
  }
};

}  // namespace b2g
