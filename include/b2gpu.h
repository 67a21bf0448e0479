/*
 * b2gpu.h — C ABI of the B200-native step engine that replaces the hot path of
 * box2d-rs `B2world::step` (reference: src/b2_world.rs:98 ->
 * src/private/dynamics/b2_world.rs:903).
 *
 * The reference has no FFI of its own (pure safe Rust).  The entry points below
 * are what a Rust `extern "C"` block in a replaced `private::step` would bind;
 * every function cites the reference interface it stands in for.  Conventions:
 *   - plain pointers and sizes only, no C++/torch types;
 *   - every call returns 0 on success or a negative B2GPU_E_* code, never
 *     unwinds; `b2gpu_last_error()` gives the message of the last failure on
 *     the calling thread;
 *   - a handle is used by one host thread at a time (the reference world is
 *     `Rc<RefCell<..>>`, i.e. !Send, src/b2_world.rs:19);
 *   - there is NO CPU fallback: without a CUDA device every stepping call
 *     fails with B2GPU_E_NO_DEVICE.
 *
 * Data model.  World state crosses the boundary as a *snapshot*: flat arrays of
 * fixed-layout records in creation order.  All intrusive lists of the reference
 * are push_front lists (src/b2rs_double_linked_list.rs:200-219), so iteration
 * order == descending creation order; arrays in ascending creation order carry
 * the same information.
 */
#ifndef B2GPU_H
#define B2GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2GPU_ABI_VERSION 2 /* 2: joint table in the snapshot (b2gpu_joint_rec, b2gpu_snapshot.joints) */

/* error codes */
#define B2GPU_OK 0
#define B2GPU_E_INVALID (-1)   /* bad argument / misuse (reference: panic!/b2_assert) */
#define B2GPU_E_NO_DEVICE (-2) /* no CUDA device: there is no CPU fallback */
#define B2GPU_E_CUDA (-3)      /* CUDA runtime error, see b2gpu_last_error */
#define B2GPU_E_CAPACITY (-4)  /* a device-side capacity was exceeded */
#define B2GPU_E_UNSUPPORTED (-5) /* feature outside the hot-path scope (TOI sub-stepping, an unknown joint type, an unregistered shape pair) */
#define B2GPU_E_LOCKED (-6)    /* world is locked (reference: is_locked() panic) */
#define B2GPU_E_INTERNAL (-7)  /* a device-side consistency check failed (a bug: please report) */
#define B2GPU_E_IO (-8)        /* a checkpoint file could not be opened, read or written */

/* body types: src/b2_body.rs B2bodyType */
#define B2GPU_STATIC_BODY 0
#define B2GPU_KINEMATIC_BODY 1
#define B2GPU_DYNAMIC_BODY 2

/* BodyFlags bits: src/b2_body.rs:209-221 */
#define B2GPU_BODY_ISLAND 0x0001u
#define B2GPU_BODY_AWAKE 0x0002u
#define B2GPU_BODY_AUTO_SLEEP 0x0004u
#define B2GPU_BODY_BULLET 0x0008u
#define B2GPU_BODY_FIXED_ROTATION 0x0010u
#define B2GPU_BODY_ENABLED 0x0020u
#define B2GPU_BODY_TOI 0x0040u

/* ContactFlags bits: src/b2_contact.rs:283-305 */
#define B2GPU_CONTACT_ISLAND 0x0001u
#define B2GPU_CONTACT_TOUCHING 0x0002u
#define B2GPU_CONTACT_ENABLED 0x0004u
#define B2GPU_CONTACT_FILTER 0x0008u

/* shape types: src/b2_shape.rs:25-31 */
#define B2GPU_SHAPE_CIRCLE 0
#define B2GPU_SHAPE_EDGE 1
#define B2GPU_SHAPE_POLYGON 2
#define B2GPU_SHAPE_CHAIN 3

/* manifold types: src/b2_collision.rs:63-67 */
#define B2GPU_MANIFOLD_CIRCLES 0
#define B2GPU_MANIFOLD_FACE_A 1
#define B2GPU_MANIFOLD_FACE_B 2

/* world flags */
#define B2GPU_WORLD_ALLOW_SLEEP 0x01u   /* m_allow_sleep      (b2_world.rs(private):43) */
#define B2GPU_WORLD_WARM_STARTING 0x02u /* m_warm_starting    (:37) */
#define B2GPU_WORLD_NEW_CONTACTS 0x04u  /* m_new_contacts     (:46) */
#define B2GPU_WORLD_CLEAR_FORCES 0x08u  /* m_clear_forces     (:48) */
#define B2GPU_WORLD_BLOCK_SOLVE 0x10u   /* G_BLOCK_SOLVE      (src/b2_contact.rs:25) */

/* joint types: B2jointType (src/b2_joint.rs:46-58), same numbering */
#define B2GPU_JOINT_DISTANCE 1
#define B2GPU_JOINT_FRICTION 2
#define B2GPU_JOINT_GEAR 3
#define B2GPU_JOINT_MOTOR 4
#define B2GPU_JOINT_MOUSE 5
#define B2GPU_JOINT_PRISMATIC 6
#define B2GPU_JOINT_PULLEY 7
#define B2GPU_JOINT_REVOLUTE 8
#define B2GPU_JOINT_WELD 9
#define B2GPU_JOINT_WHEEL 10
/* b2gpu_joint_rec.flags */
#define B2GPU_JOINT_COLLIDE_CONNECTED 0x1u /* B2jointDef::collide_connected */
#define B2GPU_JOINT_ENABLE_LIMIT 0x2u      /* revolute: m_enable_limit */
#define B2GPU_JOINT_ENABLE_MOTOR 0x4u      /* revolute: m_enable_motor */
#define B2GPU_JOINT_GEAR_PRISMATIC_1 0x100u /* gear: joint 1 is prismatic (else revolute) */
#define B2GPU_JOINT_GEAR_PRISMATIC_2 0x200u /* gear: joint 2 is prismatic */

#define B2GPU_NULL (-1)
#define B2GPU_MAX_POLYGON_VERTICES 8

/* ---------------------------------------------------------------- records */

/* B2body (src/b2_body.rs) — the fields the step reads or writes. 128 bytes. */
typedef struct b2gpu_body_rec {
  int32_t type;
  uint32_t flags;
  float xf_px, xf_py, xf_qs, xf_qc;                 /* m_xf */
  float lc_x, lc_y, c0_x, c0_y, c_x, c_y, a0, a;    /* m_sweep (alpha0 == 0 outside TOI) */
  float vx, vy, w;                                  /* m_linear_velocity, m_angular_velocity */
  float fx, fy, torque;                             /* m_force, m_torque */
  float mass, inv_mass, inertia, inv_inertia;       /* m_mass, m_inv_mass, m_i, m_inv_i */
  float linear_damping, angular_damping, gravity_scale, sleep_time;
  int32_t fixture_head;                             /* newest fixture (m_fixture_list head) or -1 */
  int32_t fixture_count;
  int32_t reserved[2];
} b2gpu_body_rec;

/* B2fixture (src/b2_fixture.rs). 48 bytes. */
typedef struct b2gpu_fixture_rec {
  int32_t body;
  int32_t next;        /* next older fixture of the same body, or -1 */
  int32_t shape_type;  /* B2GPU_SHAPE_* of the fixture's shape (chain keeps its own type) */
  int32_t shape_first; /* first child-shape record */
  int32_t child_count; /* get_child_count(): 1, or edges of a chain */
  int32_t proxy_first; /* first proxy record, -1 when the body is disabled */
  float density, friction, restitution, restitution_threshold;
  uint16_t category_bits, mask_bits; /* B2filter (src/b2_fixture.rs:30) */
  int16_t group_index;
  uint16_t is_sensor;
} b2gpu_fixture_rec;

/* One collision child: a circle, a polygon, or one edge (an edge shape, or child
 * `i` of a chain materialised as in b2_chain_shape.rs(private):59-79). 160 bytes. */
typedef struct b2gpu_shape_rec {
  int32_t type;      /* CIRCLE, EDGE or POLYGON */
  float radius;      /* m_radius */
  int32_t count;     /* polygon vertex count */
  int32_t one_sided; /* edge: m_one_sided */
  float cx, cy;      /* polygon m_centroid / circle m_p */
  float v[16];       /* polygon m_vertices (x,y)*8 ; edge: v0,v1,v2,v3 */
  float n[16];       /* polygon m_normals */
  int32_t reserved[2];
} b2gpu_shape_rec;

/* B2fixtureProxy (src/b2_fixture.rs). 32 bytes. */
typedef struct b2gpu_proxy_rec {
  int32_t fixture;
  int32_t child_index;
  int32_t proxy_id; /* tree node id */
  int32_t reserved;
  float aabb[4];    /* tight (swept) AABB lower.x lower.y upper.x upper.y */
} b2gpu_proxy_rec;

/* B2treeNode (src/b2_dynamic_tree.rs:11-32). 40 bytes. */
typedef struct b2gpu_tree_node_rec {
  float aabb[4];   /* fat AABB */
  int32_t parent;  /* parent, or next free node when height == -1 */
  int32_t child1, child2;
  int32_t height;  /* leaf = 0, free = -1 */
  int32_t proxy;   /* user data: index into proxies, -1 for internal nodes */
  int32_t moved;
} b2gpu_tree_node_rec;

/* B2manifold (src/b2_collision.rs:104-114). 64 bytes. */
typedef struct b2gpu_manifold_point {
  float lp_x, lp_y;       /* local_point */
  float normal_impulse, tangent_impulse;
  uint32_t id;            /* B2contactFeature: index_a | index_b<<8 | type_a<<16 | type_b<<24 */
} b2gpu_manifold_point;

typedef struct b2gpu_manifold {
  b2gpu_manifold_point points[2];
  float ln_x, ln_y; /* local_normal */
  float lp_x, lp_y; /* local_point */
  int32_t type;
  int32_t point_count;
} b2gpu_manifold;

/* B2contact (src/b2_contact.rs). 104 bytes. Array order = creation order
 * (world contact list reversed); per-body edge lists follow from it. */
typedef struct b2gpu_contact_rec {
  int32_t fixture_a, fixture_b;
  int32_t index_a, index_b; /* child indices */
  uint32_t flags;
  float friction, restitution, restitution_threshold, tangent_speed;
  int32_t reserved;
  b2gpu_manifold manifold;
} b2gpu_contact_rec;

/* B2joint + the derived joint (src/b2_joint.rs:160-180; revolute: src/joints/b2_revolute_joint.rs:104-136,
 * distance: src/joints/b2_distance_joint.rs) — definition, user-settable parameters and the accumulated impulses
 * the solver warm-starts from.  Array order = creation order; a body's joint list (push_front of edge A on body A,
 * then edge B on body B, src/private/dynamics/b2_world.rs:176-208) follows from it.  96 bytes.
 *   param[]  revolute: 0 reference_angle, 1 lower_angle, 2 upper_angle, 3 max_motor_torque, 4 motor_speed
 *            distance: 0 length, 1 min_length, 2 max_length, 3 stiffness, 4 damping
 *   impulse[] revolute: 0,1 m_impulse.xy, 2 m_motor_impulse, 3 m_lower_impulse, 4 m_upper_impulse
 *            distance: 0 m_impulse, 3 m_lower_impulse, 4 m_upper_impulse
 *   gear (four bodies; src/joints/b2_gear_joint.rs:150-200): body_a / body_b as given, local_anchor_a / _b copied from the
 *            coupled joints; param 0,1 local_anchor_c, 2,3 local_anchor_d, 4,5 local_axis_c, 6,7 local_axis_d; flags
 *            GEAR_PRISMATIC_1 / _2 = the coupled joints' types; impulse 0 m_impulse, and the static rest overflows into
 *            impulse 1 reference_angle_a, 2 reference_angle_b, 3 constant, 4 ratio, 5 / 6 body_c / body_d (int32 bits) */
typedef struct b2gpu_joint_rec {
  int32_t type;             /* B2GPU_JOINT_* */
  int32_t body_a, body_b;
  uint32_t flags;           /* B2GPU_JOINT_COLLIDE_CONNECTED | ENABLE_LIMIT | ENABLE_MOTOR */
  float local_anchor_a[2], local_anchor_b[2];
  float param[8];
  float impulse[8];
} b2gpu_joint_rec;

/* World-level scalars: B2world + B2broadPhase + B2dynamicTree bookkeeping. */
typedef struct b2gpu_world_rec {
  float gravity_x, gravity_y;
  float inv_dt0;        /* m_inv_dt0 */
  uint32_t flags;       /* B2GPU_WORLD_* */
  int32_t tree_root;    /* m_root */
  int32_t tree_free_list;
  int32_t tree_node_count;
  int32_t tree_node_capacity;
  int32_t tree_insertion_count;
  int32_t proxy_count;  /* B2broadPhase::m_proxy_count */
  int32_t reserved[2];
} b2gpu_world_rec;

typedef struct b2gpu_snapshot_sizes {
  int32_t body_count, fixture_count, shape_count, proxy_count;
  int32_t node_count; /* == tree_node_capacity: the whole pool incl. free nodes */
  int32_t contact_count, move_count;
  int32_t joint_count;
} b2gpu_snapshot_sizes;

/* Full step state.  Arrays are caller-owned.  For download the caller allocates
 * at least the sizes reported by *_snapshot_sizes and passes capacities in `n`. */
typedef struct b2gpu_snapshot {
  b2gpu_world_rec world;
  b2gpu_snapshot_sizes n;
  b2gpu_body_rec* bodies;         /* creation order; world body list = reverse */
  b2gpu_fixture_rec* fixtures;    /* creation order */
  b2gpu_shape_rec* shapes;
  b2gpu_proxy_rec* proxies;
  b2gpu_tree_node_rec* nodes;
  b2gpu_contact_rec* contacts;    /* creation order; world contact list = reverse */
  int32_t* move_buffer;           /* B2broadPhase::m_move_buffer (proxy ids, -1 = nulled) */
  b2gpu_joint_rec* joints;        /* creation order (ABI 2); may be NULL when n.joint_count == 0 */
} b2gpu_snapshot;

/* Counters of the last step of one world (parity quick-check + roofline inputs). */
typedef struct b2gpu_step_stats {
  int32_t status;        /* 0 or B2GPU_E_CAPACITY / B2GPU_E_UNSUPPORTED raised on device */
  int32_t contacts;      /* live contacts after the step */
  int32_t touching;      /* contacts with TOUCHING after collide */
  int32_t destroyed;     /* contacts destroyed in collide */
  int32_t islands;
  int32_t island_bodies; /* sum of island sizes (static bodies counted per island) */
  int32_t island_contacts;
  int32_t moved;         /* proxies re-inserted (move buffer length) */
  int32_t pairs;         /* pairs reported by update_pairs (the reference's pair buffer length) */
  int32_t created;       /* contacts created by add_pair */
  int32_t awake_bodies;
  int32_t solver_levels; /* large-world mode: dependency levels of one sweep, summed over the islands swept level by level
                            (b2gpu_world_set_level_threshold); 0 when every island took the one-thread form */
  int32_t reserved[4];
} b2gpu_step_stats;

/* --------------------------------------------------------------- definitions */

/* B2bodyDef (src/b2_body.rs:39-58) */
typedef struct b2gpu_body_def {
  int32_t type;
  float position_x, position_y, angle;
  float linear_velocity_x, linear_velocity_y, angular_velocity;
  float linear_damping, angular_damping;
  int32_t allow_sleep, awake, fixed_rotation, bullet, enabled;
  float gravity_scale;
} b2gpu_body_def;

/* B2fixtureDef (src/b2_fixture.rs:44-57) */
typedef struct b2gpu_fixture_def {
  float friction, restitution, restitution_threshold, density;
  int32_t is_sensor;
  uint16_t category_bits, mask_bits;
  int16_t group_index;
  uint16_t reserved;
} b2gpu_fixture_def;

/* A shape as user code builds it (B2circleShape / B2edgeShape / B2polygonShape /
 * B2chainShape).  Polygons carry already-computed hull data; use
 * b2gpu_polygon_set / b2gpu_polygon_set_as_box to fill them. */
typedef struct b2gpu_shape_def {
  int32_t type;
  float radius;
  /* circle */
  float p_x, p_y;
  /* edge */
  float v0[2], v1[2], v2[2], v3[2];
  int32_t one_sided;
  /* polygon */
  int32_t count;
  float centroid[2];
  float vertices[2 * B2GPU_MAX_POLYGON_VERTICES];
  float normals[2 * B2GPU_MAX_POLYGON_VERTICES];
  /* chain: caller-owned vertex array (x,y)*chain_count, plus ghost vertices */
  const float* chain_vertices;
  int32_t chain_count;
  float chain_prev[2], chain_next[2];
} b2gpu_shape_def;

typedef struct b2gpu_mass_data {
  float mass, center_x, center_y, inertia;
} b2gpu_mass_data;

/* B2revoluteJointDef / B2distanceJointDef (src/joints/b2_revolute_joint.rs:10-72, b2_distance_joint.rs:11-58) as one
 * plain struct; fill it with b2gpu_revolute_joint_def / b2gpu_distance_joint_def (the reference's Default + initialize)
 * and edit the fields before b2gpu_world_create_joint. */
typedef struct b2gpu_joint_def {
  int32_t type;
  int32_t body_a, body_b;
  int32_t collide_connected;
  float local_anchor_a[2], local_anchor_b[2];
  /* revolute */
  float reference_angle, lower_angle, upper_angle, max_motor_torque, motor_speed;
  int32_t enable_limit, enable_motor;
  /* distance */
  float length, min_length, max_length, stiffness, damping;
} b2gpu_joint_def;

/* Device-side capacities of one world of a batch. 0 = derive from the prototype. */
typedef struct b2gpu_caps {
  int32_t max_bodies, max_fixtures, max_shapes, max_proxies; /* reserved: topology is fixed by the prototype */
  int32_t max_contacts; /* contacts per world (default: 10 per proxy for batches, 40 for a single world) */
  int32_t max_pairs;    /* ignored: pairs are handed to add_pair as the tree query reports them, no pair buffer */
  int32_t reserved[2]; /* reserved[0]: worlds per memory block (power of two; 0 = 32 for >= 32 worlds, else 1);
                          reserved[1]: diagnostic switch, 0 = default kernels.  1 generic global-memory solver stages,
                          5 one stream (no stream groups), 6 no CUDA graphs (all bit-identical);
                          11 / 12 large-world mode (one world only; see b2gpu_world_set_large_mode flag 1 / 2) */
} b2gpu_caps;

typedef struct b2gpu_ctx b2gpu_ctx;
typedef struct b2gpu_world b2gpu_world;
typedef struct b2gpu_batch b2gpu_batch;

/* ------------------------------------------------------------------ library */
int b2gpu_abi_version(void);
const char* b2gpu_last_error(void);
/* Number of CUDA devices visible, or a negative error code. */
int b2gpu_device_count(void);
/* One context per device; creates the streams.  `stream` may be a caller's
 * cudaStream_t (e.g. torch's current stream) or NULL for a private stream. */
int b2gpu_init(int device, void* stream, b2gpu_ctx** out);
void b2gpu_shutdown(b2gpu_ctx* ctx);
int b2gpu_sync(b2gpu_ctx* ctx);
void* b2gpu_stream(b2gpu_ctx* ctx);
/* Total kernel launches issued by this context so far (bench `gpu_launches`). */
int64_t b2gpu_launch_count(b2gpu_ctx* ctx);

/* ----------------------------------------------- shapes (setup-time geometry) */
/* B2polygonShape::set_as_box (b2_polygon_shape.rs(private):15-27) */
int b2gpu_polygon_set_as_box(b2gpu_shape_def* s, float hx, float hy);
/* B2polygonShape::set_as_box_angle (:29-57) */
int b2gpu_polygon_set_as_box_angle(b2gpu_shape_def* s, float hx, float hy, float cx, float cy, float angle);
/* B2polygonShape::set (:96-211): weld, gift-wrap hull, normals, centroid */
int b2gpu_polygon_set(b2gpu_shape_def* s, const float* vertices_xy, int count);
/* B2shapeDynTrait::compute_mass for circle/edge/polygon/chain */
int b2gpu_shape_compute_mass(const b2gpu_shape_def* s, float density, b2gpu_mass_data* out);

/* ------------------------------------------- single world (B2world mirror) */
/* B2world::new(gravity) (src/b2_world.rs:22; private :26-56).  continuous_physics
 * is fixed to false (TOI is out of scope, BASELINE.json north_star). */
int b2gpu_world_create(b2gpu_ctx* ctx, float gravity_x, float gravity_y, b2gpu_world** out);
void b2gpu_world_destroy(b2gpu_world* w);
/* B2world::create_body (private :83-98). Returns the body index (>= 0). */
int b2gpu_world_create_body(b2gpu_world* w, const b2gpu_body_def* def);
/* B2body::create_fixture (b2_body.rs(private):155-200). Returns the fixture index. */
int b2gpu_body_create_fixture(b2gpu_world* w, int body, const b2gpu_fixture_def* def, const b2gpu_shape_def* shape);
/* B2body::set_transform (b2_body.rs(private):418-444) */
int b2gpu_body_set_transform(b2gpu_world* w, int body, float px, float py, float angle);
/* B2body::set_linear_velocity / set_angular_velocity (src/b2_body.rs) */
int b2gpu_body_set_linear_velocity(b2gpu_world* w, int body, float vx, float vy);
int b2gpu_body_set_angular_velocity(b2gpu_world* w, int body, float w_);
/* B2body::apply_force_to_center(force, wake) */
int b2gpu_body_apply_force_to_center(b2gpu_world* w, int body, float fx, float fy, int wake);
/* B2body::apply_force / apply_torque / apply_linear_impulse / apply_linear_impulse_to_center /
 * apply_angular_impulse (src/b2_body.rs:869-972) and set_awake (:783-801): only dynamic bodies react, `wake` wakes a
 * sleeping body first, a body that stays asleep accumulates nothing; points are world points. */
int b2gpu_body_apply_force(b2gpu_world* w, int body, float fx, float fy, float point_x, float point_y, int wake);
int b2gpu_body_apply_torque(b2gpu_world* w, int body, float torque, int wake);
int b2gpu_body_apply_linear_impulse(b2gpu_world* w, int body, float ix, float iy, float point_x, float point_y, int wake);
int b2gpu_body_apply_linear_impulse_to_center(b2gpu_world* w, int body, float ix, float iy, int wake);
int b2gpu_body_apply_angular_impulse(b2gpu_world* w, int body, float impulse, int wake);
int b2gpu_body_set_awake(b2gpu_world* w, int body, int flag);
/* B2body::set_linear_damping / set_angular_damping / set_gravity_scale (src/b2_body.rs:755-775: plain stores) and
 * set_sleeping_allowed (:815-821: clearing it wakes the body). */
int b2gpu_body_set_damping(b2gpu_world* w, int body, float linear_damping, float angular_damping);
int b2gpu_body_set_gravity_scale(b2gpu_world* w, int body, float scale);
int b2gpu_body_set_sleeping_allowed(b2gpu_world* w, int body, int flag);
/* Joints.  B2revoluteJointDef::default + ::initialize(body_a, body_b, anchor)
 * (src/joints/b2_revolute_joint.rs:10-86): local anchors and reference angle from the bodies' current transforms. */
int b2gpu_revolute_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float anchor_x, float anchor_y);
/* B2distanceJointDef::default + ::initialize(b1, b2, anchor1, anchor2) (private b2_distance_joint.rs:26-41):
 * length = max(|anchor2 - anchor1|, linear slop), min_length = max_length = length. */
int b2gpu_distance_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float a1x, float a1y, float a2x, float a2y);
/* B2prismaticJointDef::default + ::initialize(body_a, body_b, anchor, axis) (src/joints/b2_prismatic_joint.rs:10-89).  The
 * plain def struct is shared, so a prismatic def reads some fields under other names: lower_angle / upper_angle are the
 * lower / upper TRANSLATION, max_motor_torque is the maximum motor FORCE, and (length, min_length) carry local_axis_a (x, y)
 * — b2gpu_world_create_joint normalises it as B2prismaticJoint::new does; lower > upper is B2GPU_E_INVALID (the reference
 * asserts).  The revolute setters below (motor speed, max motor torque = force, enable motor / limit, set_limits) apply. */
int b2gpu_prismatic_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float anchor_x, float anchor_y,
                              float axis_x, float axis_y);
/* B2frictionJointDef::default + ::initialize(body_a, body_b, anchor) (src/joints/b2_friction_joint.rs:9-50): top-down
 * friction between two bodies.  Def overlay: `length` is max_force, `max_motor_torque` is max_torque (both default 0). */
int b2gpu_friction_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float anchor_x, float anchor_y);
/* B2motorJointDef::default + ::initialize(body_a, body_b) (src/joints/b2_motor_joint.rs:9-59): drives body B towards an
 * offset pose in body A's frame.  Def overlay: local_anchor_a is linear_offset, reference_angle is angular_offset, `length` is
 * max_force (default 1), `max_motor_torque` is max_torque (default 1), `stiffness` is correction_factor (default 0.3). */
int b2gpu_motor_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b);
/* B2pulleyJointDef::default + ::initialize(body_a, body_b, ground_a, ground_b, anchor_a, anchor_b, ratio)
 * (src/joints/b2_pulley_joint.rs:10-78): length_a + ratio * length_b stays constant.  collide_connected defaults to 1 here.
 * Def overlay: (lower_angle, upper_angle) is ground_anchor_a, (max_motor_torque, motor_speed) is ground_anchor_b, `length` is
 * length_a, `min_length` is length_b, `max_length` is the ratio.  ratio <= FLT_EPSILON here, or 0 at create: B2GPU_E_INVALID
 * (the reference asserts). */
int b2gpu_pulley_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float ground_ax, float ground_ay,
                           float ground_bx, float ground_by, float anchor_ax, float anchor_ay, float anchor_bx, float anchor_by,
                           float ratio);
/* B2gearJointDef::default (src/joints/b2_gear_joint.rs:12-40) with joint1, joint2 (revolute or prismatic joints of this
 * world, by index) and ratio: coordinate1 + ratio * coordinate2 stays constant.  body_a / body_b are set to body B of joint 1 /
 * joint 2, the bodies B2gearJoint::new measures on.  Def overlay: `enable_limit` / `enable_motor` carry the indices joint1 /
 * joint2, `length` the ratio.  Create: a coupled joint of another type, or a body B that is not dynamic: B2GPU_E_INVALID (the
 * reference asserts).  As in the reference, destroy the gear joint before either coupled joint. */
int b2gpu_gear_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int joint1, int joint2, float ratio);
/* B2mouseJointDef::default (src/joints/b2_mouse_joint.rs:8-21) with `target`: a soft constraint that drags a point of body B
 * (the body-local point under the target at creation) towards a world target; body A is only the island link (use a static
 * body).  Def overlay: local_anchor_a is the target in WORLD coordinates, `length` is max_force (default 0), stiffness and
 * damping as named (b2gpu_linear_stiffness).  Move the target with b2gpu_joint_set_target. */
int b2gpu_mouse_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float target_x, float target_y);
/* B2wheelJointDef::default + ::initialize(body_a, body_b, anchor, axis) (src/joints/b2_wheel_joint.rs:10-90): a point of
 * body B on a line of body A, with a spring (stiffness / damping: b2gpu_linear_stiffness), translation limits and a
 * rotational motor.  Def overlay as for the prismatic joint: lower_angle / upper_angle are the translation limits,
 * (length, min_length) carry local_axis_a, which B2wheelJoint::new does NOT normalise.  The revolute setters apply. */
int b2gpu_wheel_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float anchor_x, float anchor_y,
                          float axis_x, float axis_y);
/* B2weldJointDef::default + ::initialize(body_a, body_b, anchor) (src/joints/b2_weld_joint.rs:10-50): local anchors and
 * reference angle from the bodies' current transforms; stiffness = damping = 0 (rigid). */
int b2gpu_weld_joint_def(b2gpu_world* w, b2gpu_joint_def* def, int body_a, int body_b, float anchor_x, float anchor_y);
/* b2_angular_stiffness (src/private/dynamics/b2_joint.rs:47-70): stiffness and damping of a soft weld joint. */
int b2gpu_angular_stiffness(b2gpu_world* w, float frequency_hertz, float damping_ratio, int body_a, int body_b, float* stiffness,
                            float* damping);
/* b2_linear_stiffness (src/private/dynamics/b2_joint.rs:22-45): stiffness and damping of a soft distance joint. */
int b2gpu_linear_stiffness(b2gpu_world* w, float frequency_hertz, float damping_ratio, int body_a, int body_b, float* stiffness,
                           float* damping);
/* B2world::create_joint (src/private/dynamics/b2_world.rs:156-262): returns the joint index (>= 0); contacts between
 * the two bodies are flagged for filtering when collide_connected is false.  Does not wake the bodies.
 * All ten joint types of B2jointType are inside the path; any other type value: B2GPU_E_UNSUPPORTED. */
int b2gpu_world_create_joint(b2gpu_world* w, const b2gpu_joint_def* def);
/* B2world::destroy_joint (src/private/dynamics/b2_world.rs:278-339): wakes both bodies, removes the joint (the order of the
 * others is kept) and, when it had collide_connected == false, flags the contacts between its bodies for filtering so that
 * they may start to collide.  Joints created after it move down by one index. */
int b2gpu_world_destroy_joint(b2gpu_world* w, int joint);
int b2gpu_world_get_joint_count(b2gpu_world* w);
int b2gpu_world_get_joint(b2gpu_world* w, int joint, b2gpu_joint_rec* out);
/* B2revoluteJoint::set_motor_speed / set_max_motor_torque / enable_motor / enable_limit / set_limits
 * (src/joints/b2_revolute_joint.rs): wake both bodies when the value changes, as the reference does; enable_limit and
 * set_limits also zero the limit impulses. */
int b2gpu_joint_set_motor_speed(b2gpu_world* w, int joint, float speed);
int b2gpu_joint_set_max_motor_torque(b2gpu_world* w, int joint, float torque);
int b2gpu_joint_enable_motor(b2gpu_world* w, int joint, int flag);
int b2gpu_joint_enable_limit(b2gpu_world* w, int joint, int flag);
int b2gpu_joint_set_limits(b2gpu_world* w, int joint, float lower, float upper);
/* B2mouseJoint::set_target (src/joints/b2_mouse_joint.rs:114-119): wakes body B when the target changes. */
int b2gpu_joint_set_target(b2gpu_world* w, int joint, float target_x, float target_y);
/* B2world::set_gravity / get_gravity (src/b2_world.rs:232-239): the next step integrates with it; nobody is woken. */
int b2gpu_world_set_gravity(b2gpu_world* w, float gravity_x, float gravity_y);
int b2gpu_world_get_gravity(b2gpu_world* w, float* gravity_x, float* gravity_y);
/* B2world::set_allow_sleeping / set_warm_starting / set_continuous_physics */
int b2gpu_world_set_allow_sleeping(b2gpu_world* w, int flag);
int b2gpu_world_set_warm_starting(b2gpu_world* w, int flag);
int b2gpu_world_set_continuous_physics(b2gpu_world* w, int flag); /* flag != 0 -> B2GPU_E_UNSUPPORTED */
/* G_BLOCK_SOLVE (src/b2_contact.rs:25) */
int b2gpu_world_set_block_solve(b2gpu_world* w, int flag);
/* Large-world mode for ONE world of 10^4..10^6 bodies (no reference counterpart: the reference steps a world
 * on one thread).  flag != 0: the order-dependent stages of the step run data-parallel (LBVH pair finding with
 * prefix-sum compaction, parallel contact destruction, union-find islands with one thread per island keeping
 * the reference's traversal and Gauss-Seidel order).  Every step is the reference's step of the same state —
 * the pair set, the created and destroyed contact sets and all body / manifold values are identical — but
 * contacts created within one update_pairs call are appended in LBVH order instead of the order of the
 * reference's incrementally balanced tree, so free-running trajectories may diverge from the reference after
 * a step that creates several contacts on one body (SURVEY.md H1 option ii).
 * flag == 2: the same data-parallel stages, but the replica of the reference's tree is kept (one thread re-inserts
 * the moved proxies in order) and every query walks it: contacts are created in the reference's order and
 * free-running state stays bit-identical to the reference; costs the sequential re-insertion when many proxies move.
 * Default 0: exact replica tree, ordered stages one thread per world. */
int b2gpu_world_set_large_mode(b2gpu_world* w, int flag);
/* Large-world mode, giant islands: an island with at least `contacts` contact constraints and no joints is swept by one
 * CTA in dependency-level order instead of by one thread in list order (two constraints that share no movable body
 * commute, so the results are the reference's bit for bit; levels are rebuilt with the islands).  0 = the library
 * default (1024), negative = never.  Takes effect at the next island rebuild. */
int b2gpu_world_set_level_threshold(b2gpu_world* w, int contacts);
/* B2world::step (src/b2_world.rs:98; private :903-959) */
int b2gpu_world_step(b2gpu_world* w, float dt, int velocity_iterations, int position_iterations);
/* B2world::get_body_count / get_contact_count / get_proxy_count */
int b2gpu_world_get_body_count(b2gpu_world* w);
int b2gpu_world_get_contact_count(b2gpu_world* w);
/* Body getters (B2body::get_position/get_angle/get_linear_velocity/...): copies the record. */
int b2gpu_world_get_body(b2gpu_world* w, int body, b2gpu_body_rec* out);
int b2gpu_world_get_stats(b2gpu_world* w, b2gpu_step_stats* out);
/* Full step state in/out (teacher-forced parity; checkpoint/resume).  Upload (here, per batch world and at batch
 * creation) runs b2gpu_snapshot_validate first: an index outside its table is B2GPU_E_INVALID, never a device fault. */
int b2gpu_world_snapshot_sizes(b2gpu_world* w, b2gpu_snapshot_sizes* out);
int b2gpu_world_download(b2gpu_world* w, b2gpu_snapshot* out);
int b2gpu_world_upload(b2gpu_world* w, const b2gpu_snapshot* in);

/* Checkpoint / resume: a snapshot on disk.  The reference's serde support
 * (src/serialize/serialize_b2_world.rs:133-178) saves the world *definition*; a
 * world rebuilt from it has no contacts, no warm-start impulses and a fresh tree,
 * so it does not continue the saved run.  A snapshot file carries the whole step
 * state, and upload + step after load is bit-identical to the uninterrupted run.
 * Host-only calls (no device needed).  File: checksummed header + the seven
 * tables of b2gpu_snapshot as declared above, little-endian
 * (box2d_rs_b200/csrc/b2g_checkpoint.cu).
 *   validate:   every index of `s` stays inside its table (B2GPU_E_INVALID if not);
 *               save and load run it, so a corrupt file never reaches the device.
 *   save:       writes `path`.tmp, then renames it to `path`.
 *   file_sizes: table sizes of a file, for allocating the arrays.
 *   load:       `out->n` = capacities of the caller's arrays on entry, table sizes
 *               on return; B2GPU_E_CAPACITY if an array is too small, B2GPU_E_IO if
 *               the file cannot be opened, B2GPU_E_INVALID if it is not a snapshot,
 *               truncated or fails a checksum, B2GPU_E_UNSUPPORTED for another
 *               file or ABI version. */
int b2gpu_snapshot_validate(const b2gpu_snapshot* s);
int b2gpu_snapshot_save(const b2gpu_snapshot* s, const char* path);
int b2gpu_snapshot_file_sizes(const char* path, b2gpu_snapshot_sizes* out);
int b2gpu_snapshot_load(const char* path, b2gpu_snapshot* out);

/* Contact listener events for a device-resident step.  B2contactListener::begin_contact / end_contact
 * (src/b2_world_callbacks.rs:68-104) fire inside the reference's step (b2_contact.rs(private):201-211 from the
 * collide loop, b2_contact_manager.rs(private):24-49 for destroyed contacts); here they are derived on the host
 * from the snapshots taken before and after a step and returned in the reference's firing order, for replay to a
 * listener once the step has finished (pre_solve / post_solve mutations are out of scope).
 * `destroyed` = b2gpu_step_stats.destroyed of that step, or -1 to infer it.  Returns the number of events (it may
 * exceed `capacity`; `out` then holds the first `capacity`) or a negative error.  Host-only.
 * (box2d_rs_b200/csrc/b2g_events.cu) */
#define B2GPU_EVENT_BEGIN_CONTACT 1
#define B2GPU_EVENT_END_CONTACT 2
typedef struct b2gpu_contact_event {
  int32_t type;
  int32_t fixture_a, index_a, fixture_b, index_b; /* as in b2gpu_contact_rec */
  int32_t reserved[3];
} b2gpu_contact_event;
int b2gpu_contact_events(const b2gpu_snapshot* before, const b2gpu_snapshot* after, int destroyed,
                         b2gpu_contact_event* out, int capacity);

/* B2contactListener::post_solve (src/b2_world_callbacks.rs:94-103): the reference calls it from B2island::report
 * (b2_island_private.rs:460-487) for every contact of every solved island, in island order, with the impulses of the
 * contact's velocity constraint (`count` = the constraint's point count: 1 when the block solver dropped an
 * ill-conditioned second point).  With the step on the device the reports of the LAST step are read back afterwards, in
 * the reference's call order; returns their number (it may exceed `capacity`; `out` then holds the first `capacity`) or a
 * negative error.  pre_solve stays out of scope (it may mutate the contact between narrowphase and solver). */
typedef struct b2gpu_post_solve_event {
  int32_t fixture_a, index_a, fixture_b, index_b; /* as in b2gpu_contact_rec */
  int32_t count;
  float normal_impulses[2], tangent_impulses[2];
  int32_t reserved[3];
} b2gpu_post_solve_event;
int b2gpu_world_post_solve_events(b2gpu_world* w, b2gpu_post_solve_event* out, int capacity);
int b2gpu_batch_post_solve_events(b2gpu_batch* b, int world, b2gpu_post_solve_event* out, int capacity);

/* ------------------------------------------------------------ world queries
 * B2world::ray_cast (src/private/dynamics/b2_world.rs:1015-1049) with the "closest hit" callback
 * `|fixture, point, normal, fraction| fraction`, and B2world::query_aabb (:969-980) with a callback that
 * always continues, for n rays / boxes at once on the device: one thread per query walks the broadphase tree
 * (b2_dynamic_tree.rs:239-347; the LBVH in large-world mode 1) and runs the reference's shape ray casts
 * (b2_circle_shape.rs(private):26-62, b2_edge_shape.rs(private):39-102, b2_polygon_shape.rs(private):226-290,
 * b2_chain_shape.rs(private):86-108).  State = the world after its last step (or as built). */
typedef struct b2gpu_ray_hit {
  int32_t fixture;     /* -1: the ray hit nothing */
  int32_t child_index;
  float fraction;      /* output.fraction of the closest hit */
  float point_x, point_y;   /* (1 - fraction) * p1 + fraction * p2 */
  float normal_x, normal_y;
  int32_t reserved;
} b2gpu_ray_hit;
/* p1p2: host array [n][4] = p1.x p1.y p2.x p2.y (p1 != p2, the reference asserts it); out: host array [n]. */
int b2gpu_world_ray_cast_closest(b2gpu_world* w, const float* p1p2, int n, b2gpu_ray_hit* out);
/* aabbs: host array [n][4] = lower.x lower.y upper.x upper.y; counts[n] = proxies whose fat AABB overlaps box i
 * (may exceed max_hits: then only the first max_hits are stored); hits: host array [n][max_hits][2] =
 * (fixture, child index) in the order the reference's tree query reports them (LBVH order in large-world mode 1). */
int b2gpu_world_query_aabb(b2gpu_world* w, const float* aabbs, int n, int max_hits, int32_t* counts, int32_t* hits);
/* The same for every world of a batch: rays [n_worlds][rays_per_world][4], out [n_worlds][rays_per_world]
 * (an RL-style range sensor: one launch for all worlds). */
int b2gpu_batch_ray_cast_closest(b2gpu_batch* b, const float* p1p2, int rays_per_world, b2gpu_ray_hit* out);
/* B2world::query_aabb for every world of a batch (an RL-style region sensor): aabbs [n_worlds][boxes_per_world][4],
 * counts [n_worlds][boxes_per_world], hits [n_worlds][boxes_per_world][max_hits][2]; per box as
 * b2gpu_world_query_aabb (report order of the reference's tree query, true count returned, hits truncated). */
int b2gpu_batch_query_aabb(b2gpu_batch* b, const float* aabbs, int boxes_per_world, int max_hits, int32_t* counts, int32_t* hits);

/* ---------------------------------------- batched independent worlds (RL-style) */
/* n_worlds replicas of `proto`, one CTA per world per step. */
int b2gpu_batch_create(b2gpu_ctx* ctx, const b2gpu_snapshot* proto, int n_worlds, const b2gpu_caps* caps, b2gpu_batch** out);
void b2gpu_batch_destroy(b2gpu_batch* b);
int b2gpu_batch_world_count(b2gpu_batch* b);
/* Asynchronous on the context stream: `steps` consecutive B2world::step calls on every world. */
int b2gpu_batch_step(b2gpu_batch* b, float dt, int velocity_iterations, int position_iterations, int steps);
int b2gpu_batch_upload_world(b2gpu_batch* b, int world, const b2gpu_snapshot* in);
int b2gpu_batch_snapshot_sizes(b2gpu_batch* b, int world, b2gpu_snapshot_sizes* out);
int b2gpu_batch_download_world(b2gpu_batch* b, int world, b2gpu_snapshot* out);
int b2gpu_batch_get_stats(b2gpu_batch* b, int first_world, int count, b2gpu_step_stats* out);
/* Every world of the batch back to the state of `in` (same topology as the prototype): an RL-style reset of all
 * environments, the broadcast b2gpu_batch_create performs.  Synchronous. */
int b2gpu_batch_reset(b2gpu_batch* b, const b2gpu_snapshot* in);
/* Device-side failures.  The reference grows its tables; a batch has fixed capacities (b2gpu_caps), so a step can
 * overflow the contact table, the move buffer or an island list (B2GPU_E_CAPACITY), or meet an unregistered shape
 * pair where the reference panics (B2GPU_E_UNSUPPORTED).  The failure is recorded per world (b2gpu_step_stats.status),
 * stays set until that world is uploaded again, and is RETURNED by every call that synchronises anyway:
 * b2gpu_batch_step_host, b2gpu_batch_get_body_state, b2gpu_batch_download_world (buffers are still filled), and by
 * the b2gpu_world_* calls that read state back.  b2gpu_batch_step / b2gpu_world_step are asynchronous and return
 * before the device has run: poll b2gpu_batch_status (synchronises; 0 or the most negative status of any world). */
int b2gpu_batch_status(b2gpu_batch* b);
/* b2gpu_world_set_level_threshold for a batch created in a large-world mode (b2gpu_caps.reserved[1] = 11 / 12). */
int b2gpu_batch_set_level_threshold(b2gpu_batch* b, int contacts);
/* Per-body force/torque of every world for the next step (B2body::apply_force_to_center /
 * apply_torque without wake): host array [n_worlds][body_count][3], pinned or pageable. */
int b2gpu_batch_set_forces(b2gpu_batch* b, const float* host_fxfyt, int first_world, int count);
/* Linear velocity of one body index in every world (B2body::set_linear_velocity). */
int b2gpu_batch_set_linear_velocity(b2gpu_batch* b, int body, const float* host_vxvy, int first_world, int count);
/* Gravity per world (B2world::set_gravity in worlds first_world .. first_world + count; domain randomisation): host array
 * [count][2].  Takes effect at the next step, wakes nobody. */
int b2gpu_batch_set_gravity(b2gpu_batch* b, const float* host_gxgy, int first_world, int count);
/* Per-world joint controls of a batch — the RL action on a jointed agent.  One joint index (shared topology), one value per
 * world: host array [count] for MOTOR_SPEED / MAX_MOTOR_TORQUE (revolute, prismatic — there it is the maximum motor force —
 * and wheel joints), [count][2] for TARGET (mouse joints).  Semantics of B2revoluteJoint::set_motor_speed /
 * set_max_motor_torque (src/joints/b2_revolute_joint.rs:172-205) and B2mouseJoint::set_target
 * (src/joints/b2_mouse_joint.rs:114-119) in every world: a value that differs from the world's current one wakes the joint's
 * bodies (body B only for TARGET) and replaces it.  Another joint type: B2GPU_E_INVALID. */
#define B2GPU_JOINT_CONTROL_MOTOR_SPEED 0
#define B2GPU_JOINT_CONTROL_MAX_MOTOR_TORQUE 1
#define B2GPU_JOINT_CONTROL_TARGET 2
int b2gpu_batch_set_joint_control(b2gpu_batch* b, int joint, int control, const float* host_values, int first_world, int count);
/* Body state of every world after the last step: host array [count][body_count][8] =
 * (c.x, c.y, a, v.x, v.y, w, xf.p.x, xf.p.y) — what get_world_center/get_angle/
 * get_linear_velocity/get_angular_velocity/get_position return. */
int b2gpu_batch_get_body_state(b2gpu_batch* b, float* host_out, int first_world, int count);
/* Device buffers for zero-copy use from torch (size in bytes returned through *bytes): forces [n_worlds][body_count][3]
 * written by the caller, state [n_worlds][body_count][8] as in b2gpu_batch_get_body_state.  They are STAGING buffers:
 * b2gpu_batch_apply_device_forces adds the force buffer to the worlds (like b2gpu_batch_set_forces, without the host
 * copy) and b2gpu_batch_refresh_device_state fills the state buffer from the worlds; both are asynchronous on the context
 * stream, so  write forces -> apply -> b2gpu_batch_step -> refresh -> read state  needs no host synchronisation when the
 * caller works on that stream (b2gpu_stream). */
void* b2gpu_batch_body_state_device(b2gpu_batch* b, int64_t* bytes);
void* b2gpu_batch_forces_device(b2gpu_batch* b, int64_t* bytes);
int b2gpu_batch_apply_device_forces(b2gpu_batch* b);
int b2gpu_batch_refresh_device_state(b2gpu_batch* b);
/* One end-to-end step through HOST buffers: H2D forces, `steps` steps, D2H body state, synchronous. */
int b2gpu_batch_step_host(b2gpu_batch* b, const float* host_forces, float* host_state_out, float dt,
                          int velocity_iterations, int position_iterations, int steps);
/* The same round trip with compact I/O: only the prototype's DYNAMIC bodies, in body order (static bodies never move and
 * ignore forces).  b2gpu_batch_dynamic_bodies returns their count nd (and fills up to `capacity` body indices);
 * host_forces is [n_worlds][nd][3] (fx, fy, torque; may be NULL), host_state_out is [n_worlds][nd][6] =
 * (c.x, c.y, a, v.x, v.y, w) — 24 B per body instead of 32 B for every body (xf.p follows from c, a and the local
 * centre: src/b2_body.rs:974-977). */
int b2gpu_batch_dynamic_bodies(b2gpu_batch* b, int32_t* out_body_indices, int capacity);
int b2gpu_batch_step_host_dynamic(b2gpu_batch* b, const float* host_forces, float* host_state_out, float dt,
                                  int velocity_iterations, int position_iterations, int steps);
/* Algorithmic bytes of the last step summed over all worlds (SURVEY.md §8d formula). */
int64_t b2gpu_batch_algorithmic_bytes(b2gpu_batch* b);

/* Per-stage device timing (bench.py roofline): when on, every kernel launch of the context is
 * bracketed by a CUDA event pair on the context's stream.  get_stage_times synchronises and returns
 * the accumulated milliseconds and launch counts per stage since profiling was switched on. */
int b2gpu_stage_count(void);
const char* b2gpu_stage_name(int stage);
int b2gpu_set_profiling(b2gpu_ctx* ctx, int on);
int b2gpu_get_stage_times(b2gpu_ctx* ctx, double* ms_out, int64_t* launches_out, int n);

/* Diagnostic: sin/cos of n angles evaluated by the device's B2Rot::set path (src/b2_math.rs:372-376),
 * host buffers in and out.  Used by the tests to pin the device trigonometry against libm. */
int b2gpu_debug_sincos(b2gpu_ctx* ctx, const float* host_in, float* host_sin, float* host_cos, int n);

#ifdef __cplusplus
}
#endif
#endif /* B2GPU_H */
