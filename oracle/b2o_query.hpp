// b2o_query.hpp — TEST INFRASTRUCTURE: CPU restatement of the reference's world queries.
// Only tests/ may use it, as the checker of the device path.  Parity unpinned: the reference has no test or
// golden vector for these functions; the restatement follows the Rust text operation for operation.
//
//   B2world::query_aabb   src/private/dynamics/b2_world.rs:969-980  -> B2dynamicTree::query  src/b2_dynamic_tree.rs:239-267
//   B2world::ray_cast     src/private/dynamics/b2_world.rs:1015-1049 -> B2dynamicTree::ray_cast src/b2_dynamic_tree.rs:269-347
//   B2fixture::ray_cast   src/b2_fixture.rs:228 -> shape ray casts:
//     circle  src/private/collision/b2_circle_shape.rs:26-62
//     edge    src/private/collision/b2_edge_shape.rs:39-102
//     polygon src/private/collision/b2_polygon_shape.rs:226-290   (lower starts at 0.032 in the Rust port, :241)
//     chain   src/private/collision/b2_chain_shape.rs:86-108      (child edge is a default, two-sided edge)
#pragma once
#include <cmath>
#include <utility>
#include <vector>

#include "b2o_world.hpp"

namespace b2o {

struct RayCastInput { Vec2 p1, p2; float max_fraction; };
struct RayCastOutput { Vec2 normal; float fraction = 0.0f; };

inline bool ray_cast_circle(const Shape& s, RayCastOutput& out, const RayCastInput& in, const Transform& xf) {
  Vec2 position = xf.p + b2_mul_rot(xf.q, s.p);
  Vec2 sv = in.p1 - position;
  float b = b2_dot(sv, sv) - s.radius * s.radius;
  Vec2 r = in.p2 - in.p1;
  float c = b2_dot(sv, r);
  float rr = b2_dot(r, r);
  float sigma = c * c - rr * b;
  if (sigma < 0.0f || rr < EPSILON) return false;
  float a = -(c + sqrtf(sigma));
  if (0.0f <= a && a <= in.max_fraction * rr) {
    a /= rr;
    out.fraction = a;
    out.normal = sv + a * r;
    out.normal.normalize();
    return true;
  }
  return false;
}
inline bool ray_cast_edge(Vec2 v1, Vec2 v2, bool one_sided, RayCastOutput& out, const RayCastInput& in, const Transform& xf) {
  Vec2 p1 = b2_mul_t_rot(xf.q, in.p1 - xf.p);
  Vec2 p2 = b2_mul_t_rot(xf.q, in.p2 - xf.p);
  Vec2 d = p2 - p1;
  Vec2 e = v2 - v1;
  Vec2 normal(e.y, -e.x);
  normal.normalize();
  float numerator = b2_dot(normal, v1 - p1);
  if (one_sided && numerator > 0.0f) return false;
  float denominator = b2_dot(normal, d);
  if (denominator == 0.0f) return false;
  float t = numerator / denominator;
  if (t < 0.0f || in.max_fraction < t) return false;
  Vec2 q = p1 + t * d;
  Vec2 r = v2 - v1;
  float rr = b2_dot(r, r);
  if (rr == 0.0f) return false;
  float s = b2_dot(q - v1, r) / rr;
  if (s < 0.0f || 1.0f < s) return false;
  out.fraction = t;
  if (numerator > 0.0f) out.normal = -b2_mul_rot(xf.q, normal);
  else out.normal = b2_mul_rot(xf.q, normal);
  return true;
}
inline bool ray_cast_polygon(const Shape& s, RayCastOutput& out, const RayCastInput& in, const Transform& xf) {
  Vec2 p1 = b2_mul_t_rot(xf.q, in.p1 - xf.p);
  Vec2 p2 = b2_mul_t_rot(xf.q, in.p2 - xf.p);
  Vec2 d = p2 - p1;
  float lower = 0.032f, upper = in.max_fraction;  // sic: b2_polygon_shape.rs(private):241
  int index = -1;
  for (int i = 0; i < s.count; ++i) {
    float numerator = b2_dot(s.normals[i], s.vertices[i] - p1);
    float denominator = b2_dot(s.normals[i], d);
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
    out.fraction = lower;
    out.normal = b2_mul_rot(xf.q, s.normals[index]);
    return true;
  }
  return false;
}
inline bool ray_cast_fixture(const World& W, int fixture, int child, RayCastOutput& out, const RayCastInput& in) {
  const Fixture& f = W.fixtures[fixture];
  const Transform& xf = W.bodies[f.body].xf;
  const Shape& s = f.shape;
  switch (s.type) {
    case E_CIRCLE: return ray_cast_circle(s, out, in, xf);
    case E_EDGE: return ray_cast_edge(s.v1, s.v2, s.one_sided, out, in, xf);
    case E_POLYGON: return ray_cast_polygon(s, out, in, xf);
    default: {
      const int n = (int)s.chain.size();
      int i2 = child + 1;
      if (i2 == n) i2 = 0;
      return ray_cast_edge(s.chain[child], s.chain[i2], false, out, in, xf);
    }
  }
}

struct RayHit { int fixture = -1, child = 0; float fraction = 0.0f; Vec2 point, normal; };

// B2world::ray_cast with the callback `|fixture, point, normal, fraction| fraction` (closest hit).
inline RayHit world_ray_cast_closest(const World& W, Vec2 point1, Vec2 point2) {
  RayHit hit;
  const DynamicTree& tree = W.broad_phase.tree;
  Vec2 p1 = point1, p2 = point2;
  Vec2 r = p2 - p1;
  r.normalize();
  Vec2 v = b2_cross_sv(1.0f, r);
  Vec2 abs_v(fabsf(v.x), fabsf(v.y));
  float max_fraction = 1.0f;
  AABB seg;
  {
    Vec2 t = p1 + max_fraction * (p2 - p1);
    seg.lower = b2_min_v(p1, t);
    seg.upper = b2_max_v(p1, t);
  }
  std::vector<int> stack;
  stack.push_back(tree.root);
  while (!stack.empty()) {
    int id = stack.back();
    stack.pop_back();
    if (id == NULL_NODE) continue;
    const TreeNode& node = tree.nodes[id];
    if (!b2_test_overlap(node.aabb, seg)) continue;
    Vec2 c = node.aabb.get_center();
    Vec2 h = 0.5f * (node.aabb.upper - node.aabb.lower);
    float separation = fabsf(b2_dot(v, p1 - c)) - b2_dot(abs_v, h);
    if (separation > 0.0f) continue;
    if (node.is_leaf()) {
      RayCastInput sub = {p1, p2, max_fraction};
      const FixtureProxy& px = W.proxies[node.user_data];
      RayCastOutput out;
      float value = sub.max_fraction;
      if (ray_cast_fixture(W, px.fixture, px.child_index, out, sub)) {
        float fraction = out.fraction;
        hit.fixture = px.fixture;
        hit.child = px.child_index;
        hit.fraction = fraction;
        hit.point = (1.0f - fraction) * sub.p1 + fraction * sub.p2;
        hit.normal = out.normal;
        value = fraction;
      }
      if (value == 0.0f) return hit;
      if (value > 0.0f) {
        max_fraction = value;
        Vec2 t = p1 + max_fraction * (p2 - p1);
        seg.lower = b2_min_v(p1, t);
        seg.upper = b2_max_v(p1, t);
      }
    } else {
      stack.push_back(node.child1);
      stack.push_back(node.child2);
    }
  }
  return hit;
}

// B2world::query_aabb with a callback that always continues: (fixture, child) of every proxy reported, in order.
inline void world_query_aabb(const World& W, const AABB& box, std::vector<std::pair<int, int>>& out) {
  W.broad_phase.tree.query(
      [&](int id) {
        const FixtureProxy& px = W.proxies[W.broad_phase.tree.nodes[id].user_data];
        out.push_back(std::make_pair(px.fixture, px.child_index));
        return true;
      },
      box);
}

}  // namespace b2o
