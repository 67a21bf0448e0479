// b2g_world.cu — host-side mirror of the reference's world-building API behind the C ABI:
// B2world::new/create_body, B2body::create_fixture/set_transform/set_*_velocity/apply_force,
// B2polygonShape::set/set_as_box, compute_mass.  This is setup-time bookkeeping (SURVEY.md §2:
// "API stays on host, only mirrored into SoA"): it produces the snapshot that the device engine
// steps.  Stepping itself always runs on the GPU (a one-world batch); nothing here simulates.
//
// Reference: src/private/dynamics/b2_world.rs:26-98, b2_body.rs(private):15-90,155-200,292-350,
// 418-444, b2_fixture.rs(private):40-173, b2_polygon_shape.rs(private):15-211,314-390,
// b2_circle_shape.rs(private):79-86, b2_edge_shape.rs(private):126-132, b2_chain_shape.rs(private):59-79,135-141,
// b2_dynamic_tree.rs(private):81-168, b2_broad_phase.rs(private):33-75.
#include <string.h>

#include <new>
#include <vector>

#include "b2g_runtime.h"

using namespace b2g;

namespace {

// user shapes are kept per fixture for reset_mass_data (compute_mass needs the whole shape)
struct WorldDefs {
  std::vector<b2gpu_shape_def> shape_defs;
  std::vector<std::vector<float>> chain_storage;
};

struct HostWorld {
  b2gpu_world_rec world;
  std::vector<b2gpu_body_rec> bodies;
  std::vector<b2gpu_fixture_rec> fixtures;
  std::vector<b2gpu_shape_rec> shapes;
  std::vector<b2gpu_proxy_rec> proxies;
  std::vector<b2gpu_contact_rec> contacts;
  std::vector<b2gpu_joint_rec> joints;  // creation order
  std::vector<int> move_buffer;
  // replica tree on the host (stride-1 arrays driven by the same Tree code as the device)
  std::vector<float4> n_aabb;
  std::vector<int4> n_link;
  std::vector<int> n_moved, n_proxy;
  int ws[WS_COUNT];
  int status = 0;
};

void tree_reserve(HostWorld& w, int cap) {
  w.n_aabb.resize(cap, make_float4(0, 0, 0, 0));
  w.n_link.resize(cap, make_int4(-1, -1, -1, -1));
  w.n_moved.resize(cap, 0);
  w.n_proxy.resize(cap, -1);
}
Tree host_tree(HostWorld& w) {
  Tree t;
  t.aabb = w.n_aabb.data();
  t.link = w.n_link.data();
  t.moved = w.n_moved.data();
  t.stride = 1;
  t.ws = w.ws;
  t.ws_stride = 1;
  t.phys_cap = (int)w.n_aabb.size();
  t.status = &w.status;
  return t;
}
void host_world_init(HostWorld& w, float gx, float gy) {
  memset(&w.world, 0, sizeof(w.world));
  w.world.gravity_x = gx;
  w.world.gravity_y = gy;
  w.world.inv_dt0 = 0.0f;
  // b2_world.rs(private):37-48: warm starting, sleep, auto clear forces on; continuous physics is
  // outside the hot-path scope and fixed off.  G_BLOCK_SOLVE defaults to true.
  w.world.flags = B2GPU_WORLD_ALLOW_SLEEP | B2GPU_WORLD_WARM_STARTING | B2GPU_WORLD_CLEAR_FORCES | B2GPU_WORLD_BLOCK_SOLVE;
  memset(w.ws, 0, sizeof(w.ws));
  // B2dynamicTree::new (b2_dynamic_tree.rs(private):7-32): pool of 16 chained free nodes
  tree_reserve(w, 16);
  for (int i = 0; i < 15; ++i) w.n_link[i] = make_int4(i + 1, -1, -1, -1);
  w.n_link[15] = make_int4(-1, -1, -1, -1);
  w.ws[WS_TREE_ROOT] = -1;
  w.ws[WS_TREE_FREE] = 0;
  w.ws[WS_TREE_COUNT] = 0;
  w.ws[WS_TREE_CAP] = 16;
  w.ws[WS_TREE_INSERTIONS] = 0;
}

Xf body_xf(const b2gpu_body_rec& b) {
  Xf xf;
  xf.p = v2(b.xf_px, b.xf_py);
  xf.q.s = b.xf_qs;
  xf.q.c = b.xf_qc;
  return xf;
}

// B2broadPhase::create_proxy -> B2dynamicTree::create_proxy; grows the pool like allocate_node.
int bp_create_proxy(HostWorld& w, const Box& aabb, int proxy_index) {
  if (w.ws[WS_TREE_FREE] == -1) tree_reserve(w, w.ws[WS_TREE_CAP] * 2);
  Tree t = host_tree(w);
  Box fat;
  fat.lo = aabb.lo - v2(B2G_AABB_EXTENSION, B2G_AABB_EXTENSION);
  fat.hi = aabb.hi + v2(B2G_AABB_EXTENSION, B2G_AABB_EXTENSION);
  // the leaf itself may exhaust the pool and its parent needs one more node
  int id = t.allocate_node();
  if (id < 0) return id;
  t.setA(id, fat);
  t.moved[id] = 1;
  if (w.ws[WS_TREE_FREE] == -1) {
    tree_reserve(w, w.ws[WS_TREE_CAP] * 2);
    t = host_tree(w);
  }
  t.insert_leaf(id);
  w.n_proxy[id] = proxy_index;
  w.world.proxy_count += 1;
  w.move_buffer.push_back(id);
  return id;
}
// B2broadPhase::move_proxy -> B2dynamicTree::move_proxy (:109-168)
void bp_move_proxy(HostWorld& w, int id, const Box& aabb, V2 displacement) {
  Tree t = host_tree(w);
  const V2 r = v2(B2G_AABB_EXTENSION, B2G_AABB_EXTENSION);
  Box fat;
  fat.lo = aabb.lo - r;
  fat.hi = aabb.hi + r;
  const V2 d = B2G_AABB_MULTIPLIER * displacement;
  if (d.x < 0.0f) fat.lo.x += d.x; else fat.hi.x += d.x;
  if (d.y < 0.0f) fat.lo.y += d.y; else fat.hi.y += d.y;
  const Box tree_box = t.A(id);
  if (box_contains(tree_box, aabb)) {
    Box huge;
    huge.lo = fat.lo - 4.0f * r;
    huge.hi = fat.hi + 4.0f * r;
    if (box_contains(huge, tree_box)) return;
  }
  t.remove_leaf(id);
  t.setA(id, fat);
  t.insert_leaf(id);
  t.moved[id] = 1;
  w.move_buffer.push_back(id);
}

void fill_child_shape(b2gpu_shape_rec& r, const b2gpu_shape_def* s, int child) {
  memset(&r, 0, sizeof(r));
  r.radius = s->radius;
  if (s->type == B2GPU_SHAPE_CIRCLE) {
    r.type = B2GPU_SHAPE_CIRCLE;
    r.cx = s->p_x; r.cy = s->p_y;
    r.v[0] = s->p_x; r.v[1] = s->p_y;
  } else if (s->type == B2GPU_SHAPE_EDGE) {
    r.type = B2GPU_SHAPE_EDGE;
    r.one_sided = s->one_sided ? 1 : 0;
    r.v[0] = s->v0[0]; r.v[1] = s->v0[1]; r.v[2] = s->v1[0]; r.v[3] = s->v1[1];
    r.v[4] = s->v2[0]; r.v[5] = s->v2[1]; r.v[6] = s->v3[0]; r.v[7] = s->v3[1];
  } else if (s->type == B2GPU_SHAPE_POLYGON) {
    r.type = B2GPU_SHAPE_POLYGON;
    r.count = s->count;
    r.cx = s->centroid[0]; r.cy = s->centroid[1];
    memcpy(r.v, s->vertices, sizeof(r.v));
    memcpy(r.n, s->normals, sizeof(r.n));
  } else {
    // B2chainShape::get_child_edge (b2_chain_shape.rs(private):59-79): one-sided edge with ghost vertices
    const float* cv = s->chain_vertices;
    const int n = s->chain_count;
    r.type = B2GPU_SHAPE_EDGE;
    r.one_sided = 1;
    r.v[2] = cv[2 * child]; r.v[3] = cv[2 * child + 1];
    r.v[4] = cv[2 * (child + 1)]; r.v[5] = cv[2 * (child + 1) + 1];
    if (child > 0) { r.v[0] = cv[2 * (child - 1)]; r.v[1] = cv[2 * (child - 1) + 1]; }
    else { r.v[0] = s->chain_prev[0]; r.v[1] = s->chain_prev[1]; }
    if (child < n - 2) { r.v[6] = cv[2 * (child + 2)]; r.v[7] = cv[2 * (child + 2) + 1]; }
    else { r.v[6] = s->chain_next[0]; r.v[7] = s->chain_next[1]; }
  }
}

void polygon_box(b2gpu_shape_def* s, float hx, float hy) {
  s->type = B2GPU_SHAPE_POLYGON;
  s->radius = B2G_POLYGON_RADIUS;
  s->count = 4;
  const float vx[4] = {-hx, hx, hx, -hx}, vy[4] = {-hy, -hy, hy, hy};
  const float nx[4] = {0.0f, 1.0f, 0.0f, -1.0f}, ny[4] = {-1.0f, 0.0f, 1.0f, 0.0f};
  memset(s->vertices, 0, sizeof(s->vertices));
  memset(s->normals, 0, sizeof(s->normals));
  for (int i = 0; i < 4; ++i) {
    s->vertices[2 * i] = vx[i]; s->vertices[2 * i + 1] = vy[i];
    s->normals[2 * i] = nx[i]; s->normals[2 * i + 1] = ny[i];
  }
  s->centroid[0] = 0.0f;
  s->centroid[1] = 0.0f;
}

int shape_mass(const b2gpu_shape_def* s, float density, b2gpu_mass_data* md) {
  switch (s->type) {
    case B2GPU_SHAPE_CIRCLE: {
      md->mass = density * B2G_PI * s->radius * s->radius;
      md->center_x = s->p_x;
      md->center_y = s->p_y;
      md->inertia = md->mass * (0.5f * s->radius * s->radius + (s->p_x * s->p_x + s->p_y * s->p_y));
      return 0;
    }
    case B2GPU_SHAPE_EDGE: {
      md->mass = 0.0f;
      md->center_x = 0.5f * (s->v1[0] + s->v2[0]);
      md->center_y = 0.5f * (s->v1[1] + s->v2[1]);
      md->inertia = 0.0f;
      return 0;
    }
    case B2GPU_SHAPE_CHAIN: {
      md->mass = 0.0f; md->center_x = 0.0f; md->center_y = 0.0f; md->inertia = 0.0f;
      return 0;
    }
    case B2GPU_SHAPE_POLYGON: {
      if (s->count < 3 || s->count > B2G_MAX_POLY) { set_error("polygon vertex count out of range"); return B2GPU_E_INVALID; }
      V2 center = v2(0.0f, 0.0f);
      float area = 0.0f, inert = 0.0f;
      const V2 ref = v2(s->vertices[0], s->vertices[1]);
      const float k_inv3 = 1.0f / 3.0f;
      for (int i = 0; i < s->count; ++i) {
        const int j = i + 1 < s->count ? i + 1 : 0;
        const V2 e1 = v2(s->vertices[2 * i], s->vertices[2 * i + 1]) - ref;
        const V2 e2 = v2(s->vertices[2 * j], s->vertices[2 * j + 1]) - ref;
        const float d = cross(e1, e2);
        const float tri = 0.5f * d;
        area += tri;
        center = center + (tri * k_inv3) * (e1 + e2);
        const float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
        const float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
        inert += (0.25f * k_inv3 * d) * (intx2 + inty2);
      }
      md->mass = density * area;
      if (!(area > B2G_EPSILON)) { set_error("degenerate polygon (zero area)"); return B2GPU_E_INVALID; }
      center = (1.0f / area) * center;
      const V2 c = center + ref;
      md->center_x = c.x;
      md->center_y = c.y;
      md->inertia = density * inert;
      md->inertia += md->mass * (dot(c, c) - dot(center, center));
      return 0;
    }
  }
  set_error("unknown shape type");
  return B2GPU_E_INVALID;
}

}  // namespace

struct b2gpu_world {
  b2gpu_ctx* ctx = nullptr;
  HostWorld h;
  WorldDefs defs;
  BatchHost* dev = nullptr;
  bool host_dirty = true;   // host edits not on the device yet
  bool topo_dirty = true;   // bodies/fixtures changed: the device batch must be rebuilt
  bool dev_newer = false;   // the device holds the authoritative state
  int large = 0;            // 1: data-parallel large-world stages (b2g_large.h); 2: the same with the exact replica tree
  int level_min = 0;        // b2gpu_world_set_level_threshold: 0 = the library default
};

namespace {

void fill_snapshot(b2gpu_world* W, b2gpu_snapshot* s, std::vector<b2gpu_tree_node_rec>& nodes) {
  HostWorld& h = W->h;
  h.world.tree_root = h.ws[WS_TREE_ROOT];
  h.world.tree_free_list = h.ws[WS_TREE_FREE];
  h.world.tree_node_count = h.ws[WS_TREE_COUNT];
  h.world.tree_node_capacity = h.ws[WS_TREE_CAP];
  h.world.tree_insertion_count = h.ws[WS_TREE_INSERTIONS];
  const int cap = h.ws[WS_TREE_CAP];
  nodes.resize(cap);
  for (int i = 0; i < cap; ++i) {
    b2gpu_tree_node_rec& n = nodes[i];
    n.aabb[0] = h.n_aabb[i].x; n.aabb[1] = h.n_aabb[i].y; n.aabb[2] = h.n_aabb[i].z; n.aabb[3] = h.n_aabb[i].w;
    n.parent = h.n_link[i].x; n.child1 = h.n_link[i].y; n.child2 = h.n_link[i].z; n.height = h.n_link[i].w;
    n.proxy = n.height == 0 ? h.n_proxy[i] : -1;
    n.moved = h.n_moved[i];
  }
  s->world = h.world;
  s->n.body_count = (int)h.bodies.size(); s->n.fixture_count = (int)h.fixtures.size();
  s->n.shape_count = (int)h.shapes.size(); s->n.proxy_count = (int)h.proxies.size();
  s->n.node_count = cap; s->n.contact_count = (int)h.contacts.size(); s->n.move_count = (int)h.move_buffer.size();
  s->n.joint_count = (int)h.joints.size();
  s->joints = h.joints.empty() ? nullptr : h.joints.data();
  s->bodies = h.bodies.data(); s->fixtures = h.fixtures.data(); s->shapes = h.shapes.data();
  s->proxies = h.proxies.data(); s->nodes = nodes.data(); s->contacts = h.contacts.data();
  s->move_buffer = h.move_buffer.data();
}

// Pull the device state back into the host mirror before the host reads or edits it.
int ensure_host(b2gpu_world* W) {
  if (!W->dev_newer) return 0;
  HostWorld& h = W->h;
  b2gpu_snapshot_sizes n;
  int rc = batch_snapshot_sizes(W->dev, 0, &n);
  if (rc) return rc;
  std::vector<b2gpu_tree_node_rec> nodes(std::max(n.node_count, 1));
  h.contacts.resize(std::max(n.contact_count, 1));
  h.move_buffer.resize(std::max(n.move_count, 1));
  b2gpu_snapshot s;
  memset(&s, 0, sizeof(s));
  s.n = n;
  s.n.contact_count = (int)h.contacts.size();
  s.n.move_count = (int)h.move_buffer.size();
  s.n.node_count = (int)nodes.size();
  s.bodies = h.bodies.data(); s.fixtures = h.fixtures.data(); s.shapes = h.shapes.data(); s.proxies = h.proxies.data();
  s.nodes = nodes.data(); s.contacts = h.contacts.data(); s.move_buffer = h.move_buffer.data();
  s.n.joint_count = (int)h.joints.size();
  s.joints = h.joints.empty() ? nullptr : h.joints.data();
  rc = batch_download_world(W->dev, 0, &s);
  if (rc) return rc;
  h.contacts.resize(s.n.contact_count);
  h.move_buffer.resize(s.n.move_count);
  h.world = s.world;
  h.ws[WS_TREE_ROOT] = s.world.tree_root; h.ws[WS_TREE_FREE] = s.world.tree_free_list;
  h.ws[WS_TREE_COUNT] = s.world.tree_node_count; h.ws[WS_TREE_CAP] = s.world.tree_node_capacity;
  h.ws[WS_TREE_INSERTIONS] = s.world.tree_insertion_count;
  if ((int)h.n_aabb.size() < s.n.node_count) tree_reserve(h, s.n.node_count);
  for (int i = 0; i < s.n.node_count; ++i) {
    const b2gpu_tree_node_rec& nd = nodes[i];
    h.n_aabb[i] = make_float4(nd.aabb[0], nd.aabb[1], nd.aabb[2], nd.aabb[3]);
    h.n_link[i] = make_int4(nd.parent, nd.child1, nd.child2, nd.height);
    h.n_moved[i] = nd.moved;
  }
  W->dev_newer = false;
  // a world that overflowed a device table during a step (B2GPU_E_CAPACITY) or met an unregistered shape pair is no
  // longer the reference's world: every call that reads the state back reports it instead of returning 0
  return batch_last_download_status(W->dev);
}

int check_body(b2gpu_world* W, int body) {
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  if (body < 0 || body >= (int)W->h.bodies.size()) { set_error("body index out of range"); return B2GPU_E_INVALID; }
  return 0;
}

void set_awake(b2gpu_body_rec& b, bool flag) {  // src/b2_body.rs:783-801
  if (b.type == B2GPU_STATIC_BODY) return;
  if (flag) {
    b.flags |= B2GPU_BODY_AWAKE;
    b.sleep_time = 0.0f;
  } else {
    b.flags &= ~B2GPU_BODY_AWAKE;
    b.sleep_time = 0.0f;
    b.vx = b.vy = b.w = 0.0f;
    b.fx = b.fy = b.torque = 0.0f;
  }
}

int fixture_mass(const HostWorld& h, const std::vector<b2gpu_shape_def>& defs, int f, b2gpu_mass_data* md) {
  return shape_mass(&defs[f], h.fixtures[f].density, md);
}

}  // namespace

static WorldDefs* defs_of(b2gpu_world* W, bool) { return &W->defs; }

static int reset_mass_data(b2gpu_world* W, int bi) {  // b2_body.rs(private):292-350
  HostWorld& h = W->h;
  WorldDefs* D = defs_of(W, true);
  b2gpu_body_rec& b = h.bodies[bi];
  b.mass = 0.0f; b.inv_mass = 0.0f; b.inertia = 0.0f; b.inv_inertia = 0.0f;
  b.lc_x = 0.0f; b.lc_y = 0.0f;
  if (b.type == B2GPU_STATIC_BODY || b.type == B2GPU_KINEMATIC_BODY) {
    b.c0_x = b.xf_px; b.c0_y = b.xf_py;
    b.c_x = b.xf_px; b.c_y = b.xf_py;
    b.a0 = b.a;
    return 0;
  }
  V2 local_center = v2(0.0f, 0.0f);
  for (int f = b.fixture_head; f != -1; f = h.fixtures[f].next) {
    if (h.fixtures[f].density == 0.0f) continue;
    b2gpu_mass_data md;
    int rc = fixture_mass(h, D->shape_defs, f, &md);
    if (rc) return rc;
    b.mass += md.mass;
    local_center = local_center + md.mass * v2(md.center_x, md.center_y);
    b.inertia += md.inertia;
  }
  if (b.mass > 0.0f) {
    b.inv_mass = 1.0f / b.mass;
    local_center = b.inv_mass * local_center;
  }
  if (b.inertia > 0.0f && !(b.flags & B2GPU_BODY_FIXED_ROTATION)) {
    b.inertia -= b.mass * dot(local_center, local_center);
    b.inv_inertia = 1.0f / b.inertia;
  } else {
    b.inertia = 0.0f;
    b.inv_inertia = 0.0f;
  }
  const V2 old_center = v2(b.c_x, b.c_y);
  b.lc_x = local_center.x; b.lc_y = local_center.y;
  const V2 c0 = xf_mul(body_xf(b), local_center);
  b.c0_x = c0.x; b.c0_y = c0.y;
  b.c_x = c0.x; b.c_y = c0.y;
  const V2 dv = cross_sv(b.w, c0 - old_center);
  b.vx += dv.x;
  b.vy += dv.y;
  return 0;
}

// B2fixture::synchronize for every proxy of a fixture (b2_fixture.rs(private):147-173)
static void fixture_synchronize(HostWorld& h, int f, const Xf& xf1, const Xf& xf2) {
  const b2gpu_fixture_rec& fx = h.fixtures[f];
  if (fx.proxy_first < 0) return;
  for (int i = 0; i < fx.child_count; ++i) {
    b2gpu_proxy_rec& p = h.proxies[fx.proxy_first + i];
    const b2gpu_shape_rec* sh = &h.shapes[fx.shape_first + p.child_index];
    const Box a1 = shape_aabb(sh, xf1), a2 = shape_aabb(sh, xf2);
    const Box u = box_union(a1, a2);
    p.aabb[0] = u.lo.x; p.aabb[1] = u.lo.y; p.aabb[2] = u.hi.x; p.aabb[3] = u.hi.y;
    bp_move_proxy(h, p.proxy_id, u, box_center(a2) - box_center(a1));
  }
}

#define GUARD_BEGIN try {
#define GUARD_END                                                                       \
  }                                                                                     \
  catch (const std::bad_alloc&) { set_error("out of host memory"); return B2GPU_E_INVALID; } \
  catch (...) { set_error("unexpected C++ exception"); return B2GPU_E_INVALID; }

extern "C" {

int b2gpu_polygon_set_as_box(b2gpu_shape_def* s, float hx, float hy) {
  if (!s) { set_error("shape is NULL"); return B2GPU_E_INVALID; }
  polygon_box(s, hx, hy);
  return 0;
}
int b2gpu_polygon_set_as_box_angle(b2gpu_shape_def* s, float hx, float hy, float cx, float cy, float angle) {
  if (!s) { set_error("shape is NULL"); return B2GPU_E_INVALID; }
  polygon_box(s, hx, hy);
  s->centroid[0] = cx;
  s->centroid[1] = cy;
  Xf xf;
  xf.p = v2(cx, cy);
  xf.q = rot_from_angle(angle);
  for (int i = 0; i < 4; ++i) {
    const V2 v = xf_mul(xf, v2(s->vertices[2 * i], s->vertices[2 * i + 1]));
    const V2 n = rot_mul(xf.q, v2(s->normals[2 * i], s->normals[2 * i + 1]));
    s->vertices[2 * i] = v.x; s->vertices[2 * i + 1] = v.y;
    s->normals[2 * i] = n.x; s->normals[2 * i + 1] = n.y;
  }
  return 0;
}
int b2gpu_polygon_set(b2gpu_shape_def* s, const float* xy, int count) {
  if (!s || !xy) { set_error("polygon_set: NULL argument"); return B2GPU_E_INVALID; }
  if (count < 3 || count > B2G_MAX_POLY) {  // the reference asserts 3 <= count <= 8
    set_error("polygon_set: vertex count must be in [3, 8]");
    return B2GPU_E_INVALID;
  }
  V2 ps[B2G_MAX_POLY];
  int n = 0;
  const float weld = (0.5f * B2G_LINEAR_SLOP) * (0.5f * B2G_LINEAR_SLOP);
  for (int i = 0; i < count; ++i) {  // weld close vertices
    const V2 v = v2(xy[2 * i], xy[2 * i + 1]);
    bool unique = true;
    for (int j = 0; j < n; ++j)
      if (dist_sq(v, ps[j]) < weld) { unique = false; break; }
    if (unique) ps[n++] = v;
  }
  if (n < 3) { set_error("polygon_set: degenerate polygon"); return B2GPU_E_INVALID; }
  int i0 = 0;  // right-most point, lowest on ties
  float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    const float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
  }
  int hull[B2G_MAX_POLY];
  int m = 0, ih = i0;
  for (;;) {  // gift wrapping
    if (m >= B2G_MAX_POLY) { set_error("polygon_set: hull overflow"); return B2GPU_E_INVALID; }
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      const V2 r = ps[ie] - ps[hull[m]];
      const V2 v = ps[j] - ps[hull[m]];
      const float c = cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && dot(v, v) > dot(r, r)) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  if (m < 3) { set_error("polygon_set: degenerate polygon"); return B2GPU_E_INVALID; }
  s->type = B2GPU_SHAPE_POLYGON;
  s->radius = B2G_POLYGON_RADIUS;
  s->count = m;
  memset(s->vertices, 0, sizeof(s->vertices));
  memset(s->normals, 0, sizeof(s->normals));
  V2 vs[B2G_MAX_POLY];
  for (int i = 0; i < m; ++i) {
    vs[i] = ps[hull[i]];
    s->vertices[2 * i] = vs[i].x;
    s->vertices[2 * i + 1] = vs[i].y;
  }
  for (int i = 0; i < m; ++i) {
    const int i2 = i + 1 < m ? i + 1 : 0;
    V2 nrm = cross_vs(vs[i2] - vs[i], 1.0f);
    normalize(nrm);
    s->normals[2 * i] = nrm.x;
    s->normals[2 * i + 1] = nrm.y;
  }
  // compute_centroid (:59-94)
  V2 c = v2(0.0f, 0.0f);
  float area = 0.0f;
  const V2 origin = vs[0];
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < m; ++i) {
    const V2 p1 = vs[0] - origin, p2 = vs[i] - origin, p3 = (i + 1 < m ? vs[i + 1] : vs[0]) - origin;
    const V2 e1 = p2 - p1, e2 = p3 - p1;
    const float tri = 0.5f * cross(e1, e2);
    area += tri;
    c = c + (tri * inv3) * (p1 + p2 + p3);
  }
  if (!(area > B2G_EPSILON)) { set_error("polygon_set: zero area"); return B2GPU_E_INVALID; }
  c = (1.0f / area) * c + origin;
  s->centroid[0] = c.x;
  s->centroid[1] = c.y;
  return 0;
}
int b2gpu_shape_compute_mass(const b2gpu_shape_def* s, float density, b2gpu_mass_data* out) {
  if (!s || !out) { set_error("compute_mass: NULL argument"); return B2GPU_E_INVALID; }
  return shape_mass(s, density, out);
}

int b2gpu_world_create(b2gpu_ctx* ctx, float gx, float gy, b2gpu_world** out) {
  GUARD_BEGIN
  if (!ctx || !out) { set_error("world_create: bad argument"); return B2GPU_E_INVALID; }
  b2gpu_world* W = new b2gpu_world();
  W->ctx = ctx;
  host_world_init(W->h, gx, gy);
  *out = W;
  return 0;
  GUARD_END
}
void b2gpu_world_destroy(b2gpu_world* W) {
  if (!W) return;
  if (W->dev) batch_destroy(W->dev);
  delete W;
}

int b2gpu_world_create_body(b2gpu_world* W, const b2gpu_body_def* d) {
  GUARD_BEGIN
  if (!W || !d) { set_error("create_body: bad argument"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec b;
  memset(&b, 0, sizeof(b));
  b.type = d->type;
  if (d->bullet) b.flags |= B2GPU_BODY_BULLET;
  if (d->fixed_rotation) b.flags |= B2GPU_BODY_FIXED_ROTATION;
  if (d->allow_sleep) b.flags |= B2GPU_BODY_AUTO_SLEEP;
  if (d->awake && d->type != B2GPU_STATIC_BODY) b.flags |= B2GPU_BODY_AWAKE;
  if (d->enabled) b.flags |= B2GPU_BODY_ENABLED;
  const Rot q = rot_from_angle(d->angle);
  b.xf_px = d->position_x; b.xf_py = d->position_y; b.xf_qs = q.s; b.xf_qc = q.c;
  b.c0_x = b.c_x = d->position_x;
  b.c0_y = b.c_y = d->position_y;
  b.a0 = b.a = d->angle;
  b.vx = d->linear_velocity_x; b.vy = d->linear_velocity_y; b.w = d->angular_velocity;
  b.linear_damping = d->linear_damping; b.angular_damping = d->angular_damping; b.gravity_scale = d->gravity_scale;
  b.fixture_head = -1;
  W->h.bodies.push_back(b);
  W->host_dirty = W->topo_dirty = true;
  return (int)W->h.bodies.size() - 1;
  GUARD_END
}

int b2gpu_body_create_fixture(b2gpu_world* W, int body, const b2gpu_fixture_def* def, const b2gpu_shape_def* shape) {
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  if (!def || !shape) { set_error("create_fixture: NULL argument"); return B2GPU_E_INVALID; }
  if (shape->type < 0 || shape->type > B2GPU_SHAPE_CHAIN) { set_error("create_fixture: unknown shape type"); return B2GPU_E_INVALID; }
  if (shape->type == B2GPU_SHAPE_CHAIN && (shape->chain_count < 2 || !shape->chain_vertices)) { set_error("create_fixture: chain needs >= 2 vertices"); return B2GPU_E_INVALID; }
  if (shape->type == B2GPU_SHAPE_POLYGON && (shape->count < 3 || shape->count > B2G_MAX_POLY)) { set_error("create_fixture: polygon vertex count"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  HostWorld& h = W->h;
  WorldDefs* D = defs_of(W, true);
  if (D->shape_defs.size() != h.fixtures.size()) {
    set_error("create_fixture: fixtures cannot be added to a world restored with b2gpu_world_upload");
    return B2GPU_E_UNSUPPORTED;
  }
  b2gpu_fixture_rec fx;
  memset(&fx, 0, sizeof(fx));
  fx.body = body;
  fx.next = -1;
  fx.shape_type = shape->type;
  fx.shape_first = (int)h.shapes.size();
  fx.child_count = shape->type == B2GPU_SHAPE_CHAIN ? shape->chain_count - 1 : 1;
  fx.proxy_first = -1;
  fx.density = def->density; fx.friction = def->friction; fx.restitution = def->restitution;
  fx.restitution_threshold = def->restitution_threshold;
  fx.category_bits = def->category_bits; fx.mask_bits = def->mask_bits; fx.group_index = def->group_index;
  fx.is_sensor = def->is_sensor ? 1 : 0;
  for (int c = 0; c < fx.child_count; ++c) {
    b2gpu_shape_rec r;
    fill_child_shape(r, shape, c);
    h.shapes.push_back(r);
  }
  const int fi = (int)h.fixtures.size();
  h.fixtures.push_back(fx);
  // keep the user's shape for compute_mass (chains own a copy of their vertices)
  b2gpu_shape_def keep = *shape;
  if (shape->type == B2GPU_SHAPE_CHAIN) {
    D->chain_storage.emplace_back(shape->chain_vertices, shape->chain_vertices + 2 * shape->chain_count);
    keep.chain_vertices = D->chain_storage.back().data();
  }
  D->shape_defs.resize(fi + 1);
  D->shape_defs[fi] = keep;
  b2gpu_body_rec& b = h.bodies[body];
  if (b.flags & B2GPU_BODY_ENABLED) {  // B2fixture::create_proxies with the body's current transform
    h.fixtures[fi].proxy_first = (int)h.proxies.size();
    const Xf xf = body_xf(b);
    for (int c = 0; c < fx.child_count; ++c) {
      const Box a = shape_aabb(&h.shapes[fx.shape_first + c], xf);
      b2gpu_proxy_rec p;
      memset(&p, 0, sizeof(p));
      p.fixture = fi;
      p.child_index = c;
      p.aabb[0] = a.lo.x; p.aabb[1] = a.lo.y; p.aabb[2] = a.hi.x; p.aabb[3] = a.hi.y;
      const int pi = (int)h.proxies.size();
      p.proxy_id = bp_create_proxy(h, a, pi);
      if (p.proxy_id < 0) { set_error("tree allocation failed"); return B2GPU_E_CAPACITY; }
      h.proxies.push_back(p);
    }
  }
  h.fixtures[fi].next = b.fixture_head;  // push_front
  b.fixture_head = fi;
  b.fixture_count += 1;
  if (h.fixtures[fi].density > 0.0f) {
    rc = reset_mass_data(W, body);
    if (rc) return rc;
  }
  h.world.flags |= B2GPU_WORLD_NEW_CONTACTS;
  W->host_dirty = W->topo_dirty = true;
  return fi;
  GUARD_END
}

int b2gpu_body_set_transform(b2gpu_world* W, int body, float px, float py, float angle) {
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  rc = ensure_host(W);
  if (rc) return rc;
  HostWorld& h = W->h;
  b2gpu_body_rec& b = h.bodies[body];
  const Rot q = rot_from_angle(angle);
  b.xf_qs = q.s; b.xf_qc = q.c; b.xf_px = px; b.xf_py = py;
  const Xf xf = body_xf(b);
  const V2 c = xf_mul(xf, v2(b.lc_x, b.lc_y));
  b.c_x = c.x; b.c_y = c.y; b.a = angle;
  b.c0_x = c.x; b.c0_y = c.y; b.a0 = angle;
  for (int f = b.fixture_head; f != -1; f = h.fixtures[f].next) fixture_synchronize(h, f, xf, xf);
  h.world.flags |= B2GPU_WORLD_NEW_CONTACTS;
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_body_set_linear_velocity(b2gpu_world* W, int body, float vx, float vy) {
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec& b = W->h.bodies[body];
  if (b.type == B2GPU_STATIC_BODY) return 0;
  if (vx * vx + vy * vy > 0.0f) set_awake(b, true);
  b.vx = vx; b.vy = vy;
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_body_set_angular_velocity(b2gpu_world* W, int body, float w_) {
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec& b = W->h.bodies[body];
  if (b.type == B2GPU_STATIC_BODY) return 0;
  if (w_ * w_ > 0.0f) set_awake(b, true);
  b.w = w_;
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_body_set_damping(b2gpu_world* W, int body, float linear_damping, float angular_damping) {  // src/b2_body.rs:755-765
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec& b = W->h.bodies[body];
  b.linear_damping = linear_damping; b.angular_damping = angular_damping;
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_body_set_gravity_scale(b2gpu_world* W, int body, float scale) {  // src/b2_body.rs:771-773
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  W->h.bodies[body].gravity_scale = scale;
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_body_set_sleeping_allowed(b2gpu_world* W, int body, int flag) {  // src/b2_body.rs:815-821
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec& b = W->h.bodies[body];
  if (flag) b.flags |= B2GPU_BODY_AUTO_SLEEP;
  else { b.flags &= ~B2GPU_BODY_AUTO_SLEEP; set_awake(b, true); }
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_body_apply_force_to_center(b2gpu_world* W, int body, float fx, float fy, int wake) {
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec& b = W->h.bodies[body];
  if (b.type != B2GPU_DYNAMIC_BODY) return 0;
  if (wake && !(b.flags & B2GPU_BODY_AWAKE)) set_awake(b, true);
  if (b.flags & B2GPU_BODY_AWAKE) { b.fx += fx; b.fy += fy; }
  W->host_dirty = true;
  return 0;
  GUARD_END
}
// The rest of B2body's force / impulse API (src/b2_body.rs:869-972): host-side edits of the body record between
// steps, like apply_force_to_center above.  Expression shapes follow the reference operation for operation.
static int dynamic_body_for_apply(b2gpu_world* W, int body, int wake, b2gpu_body_rec** out) {
  *out = nullptr;
  int rc = check_body(W, body);
  if (rc) return rc;
  rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_body_rec& b = W->h.bodies[body];
  if (b.type != B2GPU_DYNAMIC_BODY) return 0;
  if (wake && !(b.flags & B2GPU_BODY_AWAKE)) set_awake(b, true);
  W->host_dirty = true;
  if (b.flags & B2GPU_BODY_AWAKE) *out = &b;  // a sleeping body accumulates nothing
  return 0;
}
int b2gpu_body_apply_force(b2gpu_world* W, int body, float fx, float fy, float px, float py, int wake) {  // :869-888
  GUARD_BEGIN
  b2gpu_body_rec* b;
  int rc = dynamic_body_for_apply(W, body, wake, &b);
  if (rc || !b) return rc;
  b->fx += fx; b->fy += fy;
  const float rx = px - b->c_x, ry = py - b->c_y;
  const float t1 = rx * fy, t2 = ry * fx;
  b->torque += t1 - t2;
  return 0;
  GUARD_END
}
int b2gpu_body_apply_torque(b2gpu_world* W, int body, float torque, int wake) {  // :905-918
  GUARD_BEGIN
  b2gpu_body_rec* b;
  int rc = dynamic_body_for_apply(W, body, wake, &b);
  if (rc || !b) return rc;
  b->torque += torque;
  return 0;
  GUARD_END
}
int b2gpu_body_apply_linear_impulse(b2gpu_world* W, int body, float ix, float iy, float px, float py, int wake) {  // :920-939
  GUARD_BEGIN
  b2gpu_body_rec* b;
  int rc = dynamic_body_for_apply(W, body, wake, &b);
  if (rc || !b) return rc;
  const float dvx = b->inv_mass * ix, dvy = b->inv_mass * iy;
  b->vx += dvx; b->vy += dvy;
  const float rx = px - b->c_x, ry = py - b->c_y;
  const float t1 = rx * iy, t2 = ry * ix;
  const float dw = b->inv_inertia * (t1 - t2);
  b->w += dw;
  return 0;
  GUARD_END
}
int b2gpu_body_apply_linear_impulse_to_center(b2gpu_world* W, int body, float ix, float iy, int wake) {  // :941-957
  GUARD_BEGIN
  b2gpu_body_rec* b;
  int rc = dynamic_body_for_apply(W, body, wake, &b);
  if (rc || !b) return rc;
  const float dvx = b->inv_mass * ix, dvy = b->inv_mass * iy;
  b->vx += dvx; b->vy += dvy;
  return 0;
  GUARD_END
}
int b2gpu_body_apply_angular_impulse(b2gpu_world* W, int body, float impulse, int wake) {  // :959-972
  GUARD_BEGIN
  b2gpu_body_rec* b;
  int rc = dynamic_body_for_apply(W, body, wake, &b);
  if (rc || !b) return rc;
  const float dw = b->inv_inertia * impulse;
  b->w += dw;
  return 0;
  GUARD_END
}
int b2gpu_body_set_awake(b2gpu_world* W, int body, int flag) {  // src/b2_body.rs:783-801
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  rc = ensure_host(W);
  if (rc) return rc;
  set_awake(W->h.bodies[body], flag != 0);
  W->host_dirty = true;
  return 0;
  GUARD_END
}
// ---------------------------------------------------------------- joints
static V2 body_local_point(const b2gpu_body_rec& b, V2 world_point) { return xf_mul_t(body_xf(b), world_point); }  // src/b2_body.rs:728-730
static void joint_def_defaults(b2gpu_joint_def* d, int type, int body_a, int body_b) {
  memset(d, 0, sizeof(*d));
  d->type = type; d->body_a = body_a; d->body_b = body_b;
  d->length = 1.0f; d->min_length = 0.0f; d->max_length = B2G_MAX_FLOAT;  // B2distanceJointDef::default
}
int b2gpu_revolute_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_REVOLUTE, body_a, body_b);
  const b2gpu_body_rec &a = W->h.bodies[body_a], &b = W->h.bodies[body_b];
  const V2 la = body_local_point(a, v2(ax, ay)), lb = body_local_point(b, v2(ax, ay));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  def->reference_angle = b.a - a.a;
  return 0;
  GUARD_END
}
int b2gpu_distance_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float a1x, float a1y, float a2x, float a2y) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_DISTANCE, body_a, body_b);
  const V2 la = body_local_point(W->h.bodies[body_a], v2(a1x, a1y)), lb = body_local_point(W->h.bodies[body_b], v2(a2x, a2y));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  const V2 d = v2(a2x, a2y) - v2(a1x, a1y);
  def->length = fmax_sel(length(d), B2G_LINEAR_SLOP);
  def->min_length = def->length;
  def->max_length = def->length;
  return 0;
  GUARD_END
}
int b2gpu_prismatic_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay, float dx, float dy) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_PRISMATIC, body_a, body_b);
  const b2gpu_body_rec &a = W->h.bodies[body_a], &b = W->h.bodies[body_b];
  const V2 la = body_local_point(a, v2(ax, ay)), lb = body_local_point(b, v2(ax, ay));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  const V2 axis = rot_mul_t(body_xf(a).q, v2(dx, dy));  // get_local_vector (src/b2_body.rs:733-735)
  def->length = axis.x; def->min_length = axis.y; def->max_length = 0.0f;  // local_axis_a (see b2gpu.h)
  def->reference_angle = b.a - a.a;
  return 0;
  GUARD_END
}
int b2gpu_friction_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_FRICTION, body_a, body_b);
  const V2 la = body_local_point(W->h.bodies[body_a], v2(ax, ay)), lb = body_local_point(W->h.bodies[body_b], v2(ax, ay));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  def->length = 0.0f; def->min_length = 0.0f; def->max_length = 0.0f;  // max_force (see b2gpu.h)
  return 0;
  GUARD_END
}
int b2gpu_pulley_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float gax, float gay, float gbx, float gby,
                           float ax, float ay, float bx, float by, float ratio) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  if (!(ratio > B2G_EPSILON)) { set_error("pulley_joint_def: ratio <= epsilon (the reference asserts)"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_PULLEY, body_a, body_b);
  def->collide_connected = 1;
  const V2 la = body_local_point(W->h.bodies[body_a], v2(ax, ay)), lb = body_local_point(W->h.bodies[body_b], v2(bx, by));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  def->lower_angle = gax; def->upper_angle = gay;       // ground_anchor_a (see b2gpu.h)
  def->max_motor_torque = gbx; def->motor_speed = gby;  // ground_anchor_b
  def->length = length(v2(ax, ay) - v2(gax, gay));      // length_a
  def->min_length = length(v2(bx, by) - v2(gbx, gby));  // length_b
  def->max_length = ratio;
  return 0;
  GUARD_END
}
static bool gear_couples(int type) { return type == B2GPU_JOINT_REVOLUTE || type == B2GPU_JOINT_PRISMATIC; }
int b2gpu_gear_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int joint1, int joint2, float ratio) {
  GUARD_BEGIN
  if (!W || !def) { set_error("gear_joint_def: bad argument"); return B2GPU_E_INVALID; }
  const int nj = (int)W->h.joints.size();
  if (joint1 < 0 || joint1 >= nj || joint2 < 0 || joint2 >= nj) { set_error("gear_joint_def: joint index out of range"); return B2GPU_E_INVALID; }
  joint_def_defaults(def, B2GPU_JOINT_GEAR, W->h.joints[joint1].body_b, W->h.joints[joint2].body_b);
  def->enable_limit = joint1; def->enable_motor = joint2;               // the coupled joints (see b2gpu.h)
  def->length = ratio; def->min_length = 0.0f; def->max_length = 0.0f;
  return 0;
  GUARD_END
}
int b2gpu_mouse_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float tx, float ty) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  joint_def_defaults(def, B2GPU_JOINT_MOUSE, body_a, body_b);
  def->local_anchor_a[0] = tx; def->local_anchor_a[1] = ty;  // the target, in world coordinates (see b2gpu.h)
  def->length = 0.0f; def->min_length = 0.0f; def->max_length = 0.0f;  // max_force
  return 0;
  GUARD_END
}
int b2gpu_motor_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_MOTOR, body_a, body_b);
  const b2gpu_body_rec &a = W->h.bodies[body_a], &b = W->h.bodies[body_b];
  const V2 lo = body_local_point(a, v2(b.xf_px, b.xf_py));  // linear_offset = body A's local point of body B's position
  def->local_anchor_a[0] = lo.x; def->local_anchor_a[1] = lo.y;
  def->reference_angle = b.a - a.a;                          // angular_offset
  def->length = 1.0f; def->min_length = 0.0f; def->max_length = 0.0f;  // max_force
  def->max_motor_torque = 1.0f;                              // max_torque
  def->stiffness = 0.3f;                                     // correction_factor
  return 0;
  GUARD_END
}
int b2gpu_wheel_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay, float dx, float dy) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_WHEEL, body_a, body_b);
  const b2gpu_body_rec &a = W->h.bodies[body_a], &b = W->h.bodies[body_b];
  const V2 la = body_local_point(a, v2(ax, ay)), lb = body_local_point(b, v2(ax, ay));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  const V2 axis = rot_mul_t(body_xf(a).q, v2(dx, dy));  // get_local_vector (src/b2_body.rs:733-735)
  def->length = axis.x; def->min_length = axis.y; def->max_length = 0.0f;  // local_axis_a (see b2gpu.h)
  return 0;
  GUARD_END
}
int b2gpu_weld_joint_def(b2gpu_world* W, b2gpu_joint_def* def, int body_a, int body_b, float ax, float ay) {
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!def) { set_error("joint def is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  joint_def_defaults(def, B2GPU_JOINT_WELD, body_a, body_b);
  const b2gpu_body_rec &a = W->h.bodies[body_a], &b = W->h.bodies[body_b];
  const V2 la = body_local_point(a, v2(ax, ay)), lb = body_local_point(b, v2(ax, ay));
  def->local_anchor_a[0] = la.x; def->local_anchor_a[1] = la.y;
  def->local_anchor_b[0] = lb.x; def->local_anchor_b[1] = lb.y;
  def->reference_angle = b.a - a.a;
  return 0;
  GUARD_END
}
int b2gpu_angular_stiffness(b2gpu_world* W, float frequency_hertz, float damping_ratio, int body_a, int body_b, float* stiffness,
                            float* damping) {  // src/private/dynamics/b2_joint.rs:47-70, B2body::get_inertia (src/b2_body.rs:708-711)
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!stiffness || !damping) { set_error("angular_stiffness: NULL output"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  const b2gpu_body_rec &a = W->h.bodies[body_a], &b = W->h.bodies[body_b];
  const float ia = a.inertia + a.mass * dot(v2(a.lc_x, a.lc_y), v2(a.lc_x, a.lc_y));
  const float ib = b.inertia + b.mass * dot(v2(b.lc_x, b.lc_y), v2(b.lc_x, b.lc_y));
  float i;
  if (ia > 0.0f && ib > 0.0f) i = ia * ib / (ia + ib);
  else if (ia > 0.0f) i = ia;
  else i = ib;
  const float omega = 2.0f * B2G_PI * frequency_hertz;
  *stiffness = i * omega * omega;
  *damping = 2.0f * i * damping_ratio * omega;
  return 0;
  GUARD_END
}
int b2gpu_linear_stiffness(b2gpu_world* W, float frequency_hertz, float damping_ratio, int body_a, int body_b, float* stiffness,
                           float* damping) {  // src/private/dynamics/b2_joint.rs:22-45
  GUARD_BEGIN
  int rc = check_body(W, body_a);
  if (!rc) rc = check_body(W, body_b);
  if (rc) return rc;
  if (!stiffness || !damping) { set_error("linear_stiffness: NULL output"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  const float mass_a = W->h.bodies[body_a].mass, mass_b = W->h.bodies[body_b].mass;
  float mass;
  if (mass_a > 0.0f && mass_b > 0.0f) mass = mass_a * mass_b / (mass_a + mass_b);
  else if (mass_a > 0.0f) mass = mass_a;
  else mass = mass_b;
  const float omega = 2.0f * B2G_PI * frequency_hertz;
  *stiffness = mass * omega * omega;
  *damping = 2.0f * mass * damping_ratio * omega;
  return 0;
  GUARD_END
}
int b2gpu_world_create_joint(b2gpu_world* W, const b2gpu_joint_def* def) {  // b2_world.rs(private):156-262
  GUARD_BEGIN
  if (!W || !def) { set_error("create_joint: bad argument"); return B2GPU_E_INVALID; }
  int rc = check_body(W, def->body_a);
  if (!rc) rc = check_body(W, def->body_b);
  if (rc) return rc;
  if (def->body_a == def->body_b) { set_error("create_joint: body_a == body_b (the reference asserts)"); return B2GPU_E_INVALID; }
  if (def->type != B2GPU_JOINT_REVOLUTE && def->type != B2GPU_JOINT_DISTANCE && def->type != B2GPU_JOINT_WELD &&
      def->type != B2GPU_JOINT_PRISMATIC && def->type != B2GPU_JOINT_WHEEL && def->type != B2GPU_JOINT_FRICTION &&
      def->type != B2GPU_JOINT_MOTOR && def->type != B2GPU_JOINT_PULLEY && def->type != B2GPU_JOINT_MOUSE &&
      def->type != B2GPU_JOINT_GEAR) {
    set_error("create_joint: unknown joint type");
    return B2GPU_E_UNSUPPORTED;
  }
  if (def->type == B2GPU_JOINT_PRISMATIC && !(def->lower_angle <= def->upper_angle)) {
    set_error("create_joint: prismatic lower translation > upper translation (the reference asserts)");
    return B2GPU_E_INVALID;
  }
  if (def->type == B2GPU_JOINT_GEAR) {  // the asserts of private joints/b2_gear_joint.rs:8-30
    const int nj = (int)W->h.joints.size(), j1 = def->enable_limit, j2 = def->enable_motor;
    if (j1 < 0 || j1 >= nj || j2 < 0 || j2 >= nj || !gear_couples(W->h.joints[j1].type) || !gear_couples(W->h.joints[j2].type)) {
      set_error("create_joint: a gear joint couples two revolute / prismatic joints of this world");
      return B2GPU_E_INVALID;
    }
    rc = ensure_host(W);
    if (rc) return rc;
    if (W->h.bodies[W->h.joints[j1].body_b].type != B2GPU_DYNAMIC_BODY || W->h.bodies[W->h.joints[j2].body_b].type != B2GPU_DYNAMIC_BODY) {
      set_error("create_joint: body B of a geared joint must be dynamic (the reference asserts)");
      return B2GPU_E_INVALID;
    }
  }
  if (def->type == B2GPU_JOINT_PULLEY && def->max_length == 0.0f) {
    set_error("create_joint: pulley ratio is 0 (the reference asserts)");
    return B2GPU_E_INVALID;
  }
  rc = ensure_host(W);
  if (rc) return rc;
  HostWorld& h = W->h;
  b2gpu_joint_rec j;
  memset(&j, 0, sizeof(j));
  j.type = def->type; j.body_a = def->body_a; j.body_b = def->body_b;
  j.flags = def->collide_connected ? B2GPU_JOINT_COLLIDE_CONNECTED : 0;
  memcpy(j.local_anchor_a, def->local_anchor_a, 8);
  memcpy(j.local_anchor_b, def->local_anchor_b, 8);
  if (def->type == B2GPU_JOINT_REVOLUTE) {  // B2revoluteJoint::new (src/joints/b2_revolute_joint.rs:253-287)
    j.param[0] = def->reference_angle; j.param[1] = def->lower_angle; j.param[2] = def->upper_angle;
    j.param[3] = def->max_motor_torque; j.param[4] = def->motor_speed;
    if (def->enable_limit) j.flags |= B2GPU_JOINT_ENABLE_LIMIT;
    if (def->enable_motor) j.flags |= B2GPU_JOINT_ENABLE_MOTOR;
  } else if (def->type == B2GPU_JOINT_PRISMATIC) {  // private joints/b2_prismatic_joint.rs:99-166
    V2 axis = v2(def->length, def->min_length);
    normalize(axis);
    j.param[0] = def->reference_angle; j.param[1] = def->lower_angle; j.param[2] = def->upper_angle;
    j.param[3] = def->max_motor_torque; j.param[4] = def->motor_speed;
    j.param[5] = axis.x; j.param[6] = axis.y;
    if (def->enable_limit) j.flags |= B2GPU_JOINT_ENABLE_LIMIT;
    if (def->enable_motor) j.flags |= B2GPU_JOINT_ENABLE_MOTOR;
  } else if (def->type == B2GPU_JOINT_FRICTION) {  // B2frictionJoint::new (src/joints/b2_friction_joint.rs:120-150)
    j.param[0] = def->length; j.param[1] = def->max_motor_torque;
  } else if (def->type == B2GPU_JOINT_PULLEY) {  // B2pulleyJoint::new (private joints/b2_pulley_joint.rs:8-44)
    j.param[0] = def->lower_angle; j.param[1] = def->upper_angle; j.param[2] = def->max_motor_torque; j.param[3] = def->motor_speed;
    j.param[4] = def->length; j.param[5] = def->min_length; j.param[6] = def->max_length;
    j.param[7] = def->length + def->max_length * def->min_length;  // constant = length_a + ratio * length_b
  } else if (def->type == B2GPU_JOINT_GEAR) {  // private joints/b2_gear_joint.rs:8-140
    // bodies A / B stay the def's (B2joint::new(&def.base)); coordinates are measured on the coupled joints' own bodies
    const b2gpu_joint_rec j1 = h.joints[def->enable_limit], j2 = h.joints[def->enable_motor];
    const float ratio = def->length;
    float coordinate[2];
    const b2gpu_joint_rec* jn[2] = {&j1, &j2};
    for (int s = 0; s < 2; ++s) {
      const b2gpu_joint_rec& c = *jn[s];
      const b2gpu_body_rec &bb = h.bodies[c.body_b], &bc = h.bodies[c.body_a];  // "A" / "B" of the gear, and C / D
      const Xf xf_b = body_xf(bb), xf_c = body_xf(bc);
      const V2 anchor_c = v2(c.local_anchor_a[0], c.local_anchor_a[1]), anchor_b = v2(c.local_anchor_b[0], c.local_anchor_b[1]);
      memcpy(s == 0 ? j.local_anchor_a : j.local_anchor_b, c.local_anchor_b, 8);
      j.param[2 * s] = anchor_c.x; j.param[2 * s + 1] = anchor_c.y;
      j.impulse[1 + s] = c.param[0];  // reference angle
      if (c.type == B2GPU_JOINT_REVOLUTE) {
        j.param[4 + 2 * s] = 0.0f; j.param[5 + 2 * s] = 0.0f;
        coordinate[s] = bb.a - bc.a - c.param[0];
      } else {
        const V2 axis = v2(c.param[5], c.param[6]);
        j.param[4 + 2 * s] = axis.x; j.param[5 + 2 * s] = axis.y;
        j.flags |= s == 0 ? B2GPU_JOINT_GEAR_PRISMATIC_1 : B2GPU_JOINT_GEAR_PRISMATIC_2;
        const V2 p = rot_mul_t(xf_c.q, rot_mul(xf_b.q, anchor_b) + (xf_b.p - xf_c.p));
        coordinate[s] = dot(p - anchor_c, axis);
      }
    }
    j.impulse[3] = coordinate[0] + ratio * coordinate[1];  // constant
    j.impulse[4] = ratio;
    const int32_t body_c = j1.body_a, body_d = j2.body_a;
    memcpy(&j.impulse[5], &body_c, 4);
    memcpy(&j.impulse[6], &body_d, 4);
  } else if (def->type == B2GPU_JOINT_MOUSE) {  // B2mouseJoint::new (src/joints/b2_mouse_joint.rs:140-170)
    const V2 target = v2(def->local_anchor_a[0], def->local_anchor_a[1]);
    const V2 lb = body_local_point(h.bodies[def->body_b], target);
    j.local_anchor_a[0] = 0.0f; j.local_anchor_a[1] = 0.0f;
    j.local_anchor_b[0] = lb.x; j.local_anchor_b[1] = lb.y;
    j.param[0] = def->length; j.param[1] = def->stiffness; j.param[2] = def->damping;
    j.param[3] = target.x; j.param[4] = target.y;  // per world (j_s1), like the motor settings of a revolute joint
  } else if (def->type == B2GPU_JOINT_MOTOR) {  // B2motorJoint::new (src/joints/b2_motor_joint.rs:175-205)
    j.param[0] = def->length; j.param[1] = def->max_motor_torque;
    j.param[2] = def->reference_angle; j.param[3] = def->stiffness;
  } else if (def->type == B2GPU_JOINT_WHEEL) {  // B2wheelJoint::new (src/joints/b2_wheel_joint.rs:266-310): axis not normalised
    j.param[0] = def->stiffness; j.param[1] = def->lower_angle; j.param[2] = def->upper_angle;
    j.param[3] = def->max_motor_torque; j.param[4] = def->motor_speed;
    j.param[5] = def->length; j.param[6] = def->min_length; j.param[7] = def->damping;
    if (def->enable_limit) j.flags |= B2GPU_JOINT_ENABLE_LIMIT;
    if (def->enable_motor) j.flags |= B2GPU_JOINT_ENABLE_MOTOR;
  } else if (def->type == B2GPU_JOINT_WELD) {  // B2weldJoint::new (src/joints/b2_weld_joint.rs:152-185)
    j.param[0] = def->reference_angle;
    j.param[3] = def->stiffness; j.param[4] = def->damping;
  } else {  // b2_distance_joint_new (private b2_distance_joint.rs:43-78)
    const float min_length = fmax_sel(def->min_length, B2G_LINEAR_SLOP);
    j.param[0] = fmax_sel(def->length, B2G_LINEAR_SLOP);
    j.param[1] = min_length;
    j.param[2] = fmax_sel(def->max_length, min_length);
    j.param[3] = def->stiffness; j.param[4] = def->damping;
  }
  h.joints.push_back(j);
  if (!def->collide_connected) {  // flag the contacts between the two bodies for filtering
    for (b2gpu_contact_rec& c : h.contacts) {
      const int ba = h.fixtures[c.fixture_a].body, bb = h.fixtures[c.fixture_b].body;
      if ((ba == def->body_a && bb == def->body_b) || (ba == def->body_b && bb == def->body_a)) c.flags |= B2GPU_CONTACT_FILTER;
    }
  }
  W->host_dirty = W->topo_dirty = true;  // the joint table is batch topology
  return (int)h.joints.size() - 1;  // creating a joint doesn't wake the bodies
  GUARD_END
}
int b2gpu_world_get_joint_count(b2gpu_world* W) { return W ? (int)W->h.joints.size() : B2GPU_E_INVALID; }
static int check_joint(b2gpu_world* W, int joint, int type) {
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  if (joint < 0 || joint >= (int)W->h.joints.size()) { set_error("joint index out of range"); return B2GPU_E_INVALID; }
  const int jt = W->h.joints[joint].type;  // the revolute setters edit prismatic and wheel joints too (same switches; force for torque on a slider)
  if (type && jt != type && !(type == B2GPU_JOINT_REVOLUTE && (jt == B2GPU_JOINT_PRISMATIC || jt == B2GPU_JOINT_WHEEL))) {
    set_error("joint is not of the type this call edits");
    return B2GPU_E_INVALID;
  }
  return 0;
}
int b2gpu_world_destroy_joint(b2gpu_world* W, int joint) {  // b2_world.rs(private):278-339
  GUARD_BEGIN
  int rc = check_joint(W, joint, 0);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  HostWorld& h = W->h;
  const b2gpu_joint_rec j = h.joints[joint];
  set_awake(h.bodies[j.body_a], true);
  set_awake(h.bodies[j.body_b], true);
  h.joints.erase(h.joints.begin() + joint);
  if (!(j.flags & B2GPU_JOINT_COLLIDE_CONNECTED)) {
    for (b2gpu_contact_rec& c : h.contacts) {
      const int ba = h.fixtures[c.fixture_a].body, bb = h.fixtures[c.fixture_b].body;
      if ((ba == j.body_a && bb == j.body_b) || (ba == j.body_b && bb == j.body_a)) c.flags |= B2GPU_CONTACT_FILTER;
    }
  }
  W->host_dirty = W->topo_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_world_get_joint(b2gpu_world* W, int joint, b2gpu_joint_rec* out) {
  GUARD_BEGIN
  int rc = check_joint(W, joint, 0);
  if (rc) return rc;
  if (!out) { set_error("out is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  *out = W->h.joints[joint];
  return 0;
  GUARD_END
}
// B2revoluteJoint setters (src/joints/b2_revolute_joint.rs:172-242): wake both bodies when the value changes
static void joint_wake(b2gpu_world* W, const b2gpu_joint_rec& j) {
  set_awake(W->h.bodies[j.body_a], true);
  set_awake(W->h.bodies[j.body_b], true);
}
int b2gpu_joint_set_motor_speed(b2gpu_world* W, int joint, float speed) {
  GUARD_BEGIN
  int rc = check_joint(W, joint, B2GPU_JOINT_REVOLUTE);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_joint_rec& j = W->h.joints[joint];
  if (speed != j.param[4]) { joint_wake(W, j); j.param[4] = speed; W->host_dirty = true; }
  return 0;
  GUARD_END
}
int b2gpu_joint_set_target(b2gpu_world* W, int joint, float tx, float ty) {  // src/joints/b2_mouse_joint.rs:114-119
  GUARD_BEGIN
  int rc = check_joint(W, joint, B2GPU_JOINT_MOUSE);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_joint_rec& j = W->h.joints[joint];
  if (tx != j.param[3] || ty != j.param[4]) {
    set_awake(W->h.bodies[j.body_b], true);
    j.param[3] = tx; j.param[4] = ty;
    W->host_dirty = true;
  }
  return 0;
  GUARD_END
}
int b2gpu_joint_set_max_motor_torque(b2gpu_world* W, int joint, float torque) {
  GUARD_BEGIN
  int rc = check_joint(W, joint, B2GPU_JOINT_REVOLUTE);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_joint_rec& j = W->h.joints[joint];
  if (torque != j.param[3]) { joint_wake(W, j); j.param[3] = torque; W->host_dirty = true; }
  return 0;
  GUARD_END
}
int b2gpu_joint_enable_motor(b2gpu_world* W, int joint, int flag) {
  GUARD_BEGIN
  int rc = check_joint(W, joint, B2GPU_JOINT_REVOLUTE);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_joint_rec& j = W->h.joints[joint];
  if ((flag != 0) != ((j.flags & B2GPU_JOINT_ENABLE_MOTOR) != 0)) {
    joint_wake(W, j);
    j.flags = flag ? (j.flags | B2GPU_JOINT_ENABLE_MOTOR) : (j.flags & ~B2GPU_JOINT_ENABLE_MOTOR);
    W->host_dirty = true;
  }
  return 0;
  GUARD_END
}
int b2gpu_joint_enable_limit(b2gpu_world* W, int joint, int flag) {
  GUARD_BEGIN
  int rc = check_joint(W, joint, B2GPU_JOINT_REVOLUTE);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_joint_rec& j = W->h.joints[joint];
  if ((flag != 0) != ((j.flags & B2GPU_JOINT_ENABLE_LIMIT) != 0)) {
    joint_wake(W, j);
    j.flags = flag ? (j.flags | B2GPU_JOINT_ENABLE_LIMIT) : (j.flags & ~B2GPU_JOINT_ENABLE_LIMIT);
    j.impulse[3] = 0.0f; j.impulse[4] = 0.0f;
    W->host_dirty = true;
  }
  return 0;
  GUARD_END
}
int b2gpu_joint_set_limits(b2gpu_world* W, int joint, float lower, float upper) {
  GUARD_BEGIN
  int rc = check_joint(W, joint, B2GPU_JOINT_REVOLUTE);
  if (!rc) rc = ensure_host(W);
  if (rc) return rc;
  if (!(lower <= upper)) { set_error("set_limits: lower > upper (the reference asserts)"); return B2GPU_E_INVALID; }
  b2gpu_joint_rec& j = W->h.joints[joint];
  if (lower != j.param[1] || upper != j.param[2]) {
    joint_wake(W, j);
    j.impulse[3] = 0.0f; j.impulse[4] = 0.0f;
    j.param[1] = lower; j.param[2] = upper;
    W->host_dirty = W->topo_dirty = true;  // the limits are static batch topology
  }
  return 0;
  GUARD_END
}

static int set_world_flag(b2gpu_world* W, uint32_t bit, int flag) {
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  if (flag) W->h.world.flags |= bit; else W->h.world.flags &= ~bit;
  W->host_dirty = true;
  return 0;
}
int b2gpu_world_set_allow_sleeping(b2gpu_world* W, int flag) {  // b2_world.rs(private):340-353
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  const bool cur = (W->h.world.flags & B2GPU_WORLD_ALLOW_SLEEP) != 0;
  if ((flag != 0) == cur) return 0;
  rc = set_world_flag(W, B2GPU_WORLD_ALLOW_SLEEP, flag);
  if (rc) return rc;
  if (!flag)
    for (auto& b : W->h.bodies) set_awake(b, true);
  return 0;
  GUARD_END
}
int b2gpu_world_set_gravity(b2gpu_world* W, float gx, float gy) {  // src/b2_world.rs:358-360: no waking, takes effect next step
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  W->h.world.gravity_x = gx; W->h.world.gravity_y = gy;
  W->host_dirty = true;
  return 0;
  GUARD_END
}
int b2gpu_world_get_gravity(b2gpu_world* W, float* gx, float* gy) {
  if (!W || !gx || !gy) { set_error("get_gravity: bad argument"); return B2GPU_E_INVALID; }
  *gx = W->h.world.gravity_x; *gy = W->h.world.gravity_y;  // only the host changes it
  return 0;
}
int b2gpu_world_set_warm_starting(b2gpu_world* W, int flag) { return set_world_flag(W, B2GPU_WORLD_WARM_STARTING, flag); }
int b2gpu_world_set_block_solve(b2gpu_world* W, int flag) { return set_world_flag(W, B2GPU_WORLD_BLOCK_SOLVE, flag); }
int b2gpu_world_set_large_mode(b2gpu_world* W, int flag) {
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  if (flag < 0 || flag > 2) { set_error("large mode: 0, 1 or 2"); return B2GPU_E_INVALID; }
  if (flag == W->large) return 0;
  int rc = ensure_host(W);  // bring the state back before the device batch is rebuilt in the other mode
  if (rc) return rc;
  W->large = flag;
  W->topo_dirty = true;
  return 0;
}
int b2gpu_world_set_level_threshold(b2gpu_world* W, int contacts) {
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  W->level_min = contacts;
  if (W->dev) W->dev->lw_level_min = contacts == 0 ? (int)LW_LEVEL_MIN_DEFAULT : contacts;  // takes effect at the next island rebuild
  return 0;
}
int b2gpu_world_set_continuous_physics(b2gpu_world* W, int flag) {
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  if (flag) { set_error("continuous physics (TOI sub-stepping) is outside the hot-path scope"); return B2GPU_E_UNSUPPORTED; }
  return 0;
}

// The device batch of a world, created / refreshed from the host mirror when the user edited the world.
static int ensure_device(b2gpu_world* W) {
  if (W->h.bodies.empty()) { set_error("world has no bodies"); return B2GPU_E_INVALID; }
  int rc;
  if (!W->dev || W->topo_dirty) {
    rc = ensure_host(W);
    if (rc) return rc;
    if (W->dev) { batch_destroy(W->dev); W->dev = nullptr; }
    b2gpu_snapshot s;
    std::vector<b2gpu_tree_node_rec> nodes;
    fill_snapshot(W, &s, nodes);
    b2gpu_caps caps;
    memset(&caps, 0, sizeof(caps));
    caps.reserved[1] = W->large == 1 ? 11 : W->large == 2 ? 12 : 0;
    rc = batch_create(&W->ctx->c, &s, 1, &caps, 1, &W->dev);
    if (rc) return rc;
    if (W->level_min != 0) W->dev->lw_level_min = W->level_min;
  } else if (W->host_dirty) {
    b2gpu_snapshot s;
    std::vector<b2gpu_tree_node_rec> nodes;
    fill_snapshot(W, &s, nodes);
    rc = batch_upload_world(W->dev, 0, &s);
    if (rc) return rc;
  }
  W->host_dirty = W->topo_dirty = false;
  return 0;
}

int b2gpu_world_step(b2gpu_world* W, float dt, int vi, int pi) {
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_device(W);
  if (rc) return rc;
  rc = batch_step(W->dev, dt, vi, pi, 1);
  if (rc) return rc;
  W->dev_newer = true;
  return 0;
  GUARD_END
}

int b2gpu_world_ray_cast_closest(b2gpu_world* W, const float* p1p2, int n, b2gpu_ray_hit* out) {
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_device(W);
  if (rc) return rc;
  return batch_ray_cast_closest(W->dev, p1p2, n, out);
  GUARD_END
}
int b2gpu_world_query_aabb(b2gpu_world* W, const float* aabbs, int n, int max_hits, int32_t* counts, int32_t* hits) {
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_device(W);
  if (rc) return rc;
  return batch_query_aabb(W->dev, aabbs, n, max_hits, counts, hits);
  GUARD_END
}

int b2gpu_world_post_solve_events(b2gpu_world* W, b2gpu_post_solve_event* out, int capacity) {
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  if (!W->dev || W->host_dirty || W->topo_dirty) return 0;  // edited since the last step: the device tables are about to be replaced
  return batch_post_solve_events(W->dev, 0, out, capacity);
  GUARD_END
}
int b2gpu_world_get_body_count(b2gpu_world* W) { return W ? (int)W->h.bodies.size() : B2GPU_E_INVALID; }
int b2gpu_world_get_contact_count(b2gpu_world* W) {
  GUARD_BEGIN
  if (!W) { set_error("world is NULL"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  return (int)W->h.contacts.size();
  GUARD_END
}
int b2gpu_world_get_body(b2gpu_world* W, int body, b2gpu_body_rec* out) {
  GUARD_BEGIN
  int rc = check_body(W, body);
  if (rc) return rc;
  if (!out) { set_error("out is NULL"); return B2GPU_E_INVALID; }
  rc = ensure_host(W);
  if (rc) return rc;
  *out = W->h.bodies[body];
  return 0;
  GUARD_END
}
int b2gpu_world_get_stats(b2gpu_world* W, b2gpu_step_stats* out) {
  GUARD_BEGIN
  if (!W || !out) { set_error("get_stats: bad argument"); return B2GPU_E_INVALID; }
  if (!W->dev) { memset(out, 0, sizeof(*out)); return 0; }
  return batch_get_stats(W->dev, 0, 1, out);
  GUARD_END
}
int b2gpu_world_snapshot_sizes(b2gpu_world* W, b2gpu_snapshot_sizes* out) {
  GUARD_BEGIN
  if (!W || !out) { set_error("snapshot_sizes: bad argument"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  const HostWorld& h = W->h;
  out->body_count = (int)h.bodies.size(); out->fixture_count = (int)h.fixtures.size(); out->shape_count = (int)h.shapes.size();
  out->proxy_count = (int)h.proxies.size(); out->node_count = h.ws[WS_TREE_CAP]; out->contact_count = (int)h.contacts.size();
  out->move_count = (int)h.move_buffer.size(); out->joint_count = (int)h.joints.size();
  return 0;
  GUARD_END
}
int b2gpu_world_download(b2gpu_world* W, b2gpu_snapshot* out) {
  GUARD_BEGIN
  if (!W || !out) { set_error("download: bad argument"); return B2GPU_E_INVALID; }
  int rc = ensure_host(W);
  if (rc) return rc;
  b2gpu_snapshot s;
  std::vector<b2gpu_tree_node_rec> nodes;
  fill_snapshot(W, &s, nodes);
  if (out->n.body_count < s.n.body_count || out->n.fixture_count < s.n.fixture_count || out->n.shape_count < s.n.shape_count ||
      out->n.proxy_count < s.n.proxy_count || out->n.node_count < s.n.node_count || out->n.contact_count < s.n.contact_count ||
      out->n.move_count < s.n.move_count || out->n.joint_count < s.n.joint_count || (s.n.joint_count > 0 && !out->joints)) {
    set_error("download buffers smaller than snapshot_sizes");
    return B2GPU_E_INVALID;
  }
  out->world = s.world;
  out->n = s.n;
  memcpy(out->bodies, s.bodies, sizeof(b2gpu_body_rec) * s.n.body_count);
  memcpy(out->fixtures, s.fixtures, sizeof(b2gpu_fixture_rec) * s.n.fixture_count);
  memcpy(out->shapes, s.shapes, sizeof(b2gpu_shape_rec) * s.n.shape_count);
  memcpy(out->proxies, s.proxies, sizeof(b2gpu_proxy_rec) * s.n.proxy_count);
  memcpy(out->nodes, s.nodes, sizeof(b2gpu_tree_node_rec) * s.n.node_count);
  memcpy(out->contacts, s.contacts, sizeof(b2gpu_contact_rec) * s.n.contact_count);
  memcpy(out->move_buffer, s.move_buffer, sizeof(int32_t) * s.n.move_count);
  if (s.n.joint_count > 0) memcpy(out->joints, s.joints, sizeof(b2gpu_joint_rec) * s.n.joint_count);
  return 0;
  GUARD_END
}
int b2gpu_world_upload(b2gpu_world* W, const b2gpu_snapshot* in) {
  GUARD_BEGIN
  if (!W || !in) { set_error("upload: bad argument"); return B2GPU_E_INVALID; }
  HostWorld& h = W->h;
  const b2gpu_snapshot_sizes& n = in->n;
  if (n.body_count < 1 || n.node_count < 1) { set_error("upload: empty snapshot"); return B2GPU_E_INVALID; }
  { int rc = b2gpu_snapshot_validate(in); if (rc) return rc; }
  h.world = in->world;
  h.bodies.assign(in->bodies, in->bodies + n.body_count);
  h.fixtures.assign(in->fixtures, in->fixtures + n.fixture_count);
  h.shapes.assign(in->shapes, in->shapes + n.shape_count);
  h.proxies.assign(in->proxies, in->proxies + n.proxy_count);
  h.contacts.assign(in->contacts, in->contacts + n.contact_count);
  h.move_buffer.assign(in->move_buffer, in->move_buffer + n.move_count);
  h.joints.assign(in->joints, in->joints + (in->joints ? n.joint_count : 0));
  h.n_aabb.clear(); h.n_link.clear(); h.n_moved.clear(); h.n_proxy.clear();
  tree_reserve(h, n.node_count);
  for (int i = 0; i < n.node_count; ++i) {
    const b2gpu_tree_node_rec& nd = in->nodes[i];
    h.n_aabb[i] = make_float4(nd.aabb[0], nd.aabb[1], nd.aabb[2], nd.aabb[3]);
    h.n_link[i] = make_int4(nd.parent, nd.child1, nd.child2, nd.height);
    h.n_moved[i] = nd.moved;
    h.n_proxy[i] = nd.proxy;
  }
  h.ws[WS_TREE_ROOT] = in->world.tree_root; h.ws[WS_TREE_FREE] = in->world.tree_free_list;
  h.ws[WS_TREE_COUNT] = in->world.tree_node_count; h.ws[WS_TREE_CAP] = in->world.tree_node_capacity;
  h.ws[WS_TREE_INSERTIONS] = in->world.tree_insertion_count;
  W->defs.shape_defs.clear();
  W->defs.chain_storage.clear();
  W->dev_newer = false;
  W->host_dirty = W->topo_dirty = true;
  return 0;
  GUARD_END
}

}  // extern "C"
