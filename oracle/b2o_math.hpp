// TEST INFRASTRUCTURE — CPU oracle, never linked into the product (box2d_rs_b200/).
// PARITY UNPINNED beyond the reference's four relevant tests: the
// reference is Rust and no Rust toolchain exists here, so this file is a literal
// C++ restatement, compiled with -O2 -ffp-contract=off (one rounding per op, no FMA).
//
// b2o_math.hpp — restates src/b2_math.rs, src/b2_common.rs, src/b2_settings.rs.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace b2o {

// src/b2_common.rs:17-19, :25-91 ; src/b2_settings.rs:7-11
constexpr float MAX_FLOAT = FLT_MAX;
constexpr float EPSILON = FLT_EPSILON;
constexpr float PI = 3.14159265358979323846f;  // std::f32::consts::PI
constexpr float LENGTH_UNITS_PER_METER = 1.0f;
constexpr int MAX_POLYGON_VERTICES = 8;
constexpr int MAX_MANIFOLD_POINTS = 2;
constexpr float AABB_EXTENSION = 0.1f * LENGTH_UNITS_PER_METER;
constexpr float AABB_MULTIPLIER = 4.0f;
constexpr float LINEAR_SLOP = 0.005f * LENGTH_UNITS_PER_METER;
constexpr float ANGULAR_SLOP = 2.0f / 180.0f * PI;
constexpr float POLYGON_RADIUS = 2.0f * LINEAR_SLOP;
constexpr float MAX_LINEAR_CORRECTION = 0.2f * LENGTH_UNITS_PER_METER;
constexpr float MAX_ANGULAR_CORRECTION = 8.0f / 180.0f * PI;  // src/b2_common.rs:64
constexpr float MAX_TRANSLATION = 2.0f * LENGTH_UNITS_PER_METER;
constexpr float MAX_TRANSLATION_SQUARED = MAX_TRANSLATION * MAX_TRANSLATION;
constexpr float MAX_ROTATION = 0.5f * PI;
constexpr float MAX_ROTATION_SQUARED = MAX_ROTATION * MAX_ROTATION;
constexpr float BAUMGARTE = 0.2f;
constexpr float TIME_TO_SLEEP = 0.5f;
constexpr float LINEAR_SLEEP_TOLERANCE = 0.01f * LENGTH_UNITS_PER_METER;
constexpr float ANGULAR_SLEEP_TOLERANCE = 2.0f / 180.0f * PI;

// src/b2_math.rs:704-731 — compare-select, not fmin/fmax
template <class T> inline T b2_min(T a, T b) { return a < b ? a : b; }
template <class T> inline T b2_max(T a, T b) { return a > b ? a : b; }
template <class T> inline T b2_clamp(T a, T lo, T hi) { return b2_max(lo, b2_min(a, hi)); }

// src/b2_math.rs:23-97
struct Vec2 {
  float x = 0.0f, y = 0.0f;
  Vec2() = default;
  Vec2(float x_, float y_) : x(x_), y(y_) {}
  void set_zero() { x = 0.0f; y = 0.0f; }
  void set(float x_, float y_) { x = x_; y = y_; }
  float length() const { return sqrtf(x * x + y * y); }
  float length_squared() const { return x * x + y * y; }
  // :82-92 — early-out leaves the vector untouched, multiplies by the reciprocal
  float normalize() {
    float len = length();
    if (len < EPSILON) return 0.0f;
    float inv = 1.0f / len;
    x *= inv;
    y *= inv;
    return len;
  }
  Vec2 operator-() const { return Vec2(-x, -y); }
  void operator+=(Vec2 o) { x += o.x; y += o.y; }
  void operator-=(Vec2 o) { x -= o.x; y -= o.y; }
  void operator*=(float a) { x *= a; y *= a; }
};
inline Vec2 operator+(Vec2 a, Vec2 b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(Vec2 a, Vec2 b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(float s, Vec2 a) { return Vec2(s * a.x, s * a.y); }

// B2Mat33 (src/b2_math.rs:296-352, src/private/common/b2_math.rs) on a plain array: ex.xyz, ey.xyz, ez.xyz at [0..8].
inline float vec3_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }  // b2_dot_vec3 :569
inline void vec3_cross(const float* a, const float* b, float* o) {  // b2_cross_vec3 :574-580
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// get_inverse22 (private b2_math.rs:32-48): the inverse of the upper 2x2 block, the rest zero
inline void mat33_get_inverse22(const float* k, float* m) {
  float a = k[0], b = k[3], c = k[1], d = k[4];
  float det = a * d - b * c;
  if (det != 0.0f) det = 1.0f / det;
  m[0] = det * d; m[3] = -det * b; m[2] = 0.0f;
  m[1] = -det * c; m[4] = det * a; m[5] = 0.0f;
  m[6] = 0.0f; m[7] = 0.0f; m[8] = 0.0f;
}
// get_sym_inverse33 (private b2_math.rs:50-72): zero matrix if singular
inline void mat33_get_sym_inverse33(const float* k, float* m) {
  float cr[3];
  vec3_cross(k + 3, k + 6, cr);
  float det = vec3_dot(k, cr);
  if (det != 0.0f) det = 1.0f / det;
  float a11 = k[0], a12 = k[3], a13 = k[6], a22 = k[4], a23 = k[7], a33 = k[8];
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
// solve33 (private b2_math.rs:5-16)
inline void mat33_solve33(const float* k, const float* b, float* x) {
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
struct Mat22 {
  Vec2 ex, ey;
  void set_zero() { ex.set_zero(); ey.set_zero(); }
  // :261-274
  Mat22 get_inverse() const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    Mat22 m;
    m.ex = Vec2(det * d, -det * c);
    m.ey = Vec2(-det * b, det * a);
    return m;
  }
  // :276-293 — solve A * x = b
  Vec2 solve(Vec2 b) const {
    float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    return Vec2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
  }
};

// :355-376 — libm sinf/cosf, what f32::sin/cos lower to on linux-gnu
struct Rot {
  float s = 0.0f, c = 0.0f;  // derive(Default): zeros
  Rot() = default;
  explicit Rot(float angle) : s(sinf(angle)), c(cosf(angle)) {}
  void set(float angle) { s = sinf(angle); c = cosf(angle); }
};
struct Transform {
  Vec2 p;
  Rot q;
};
struct Sweep {
  Vec2 local_center, c0, c;
  float a0 = 0.0f, a = 0.0f, alpha0 = 0.0f;
};

// :470-495
inline float b2_dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
inline float b2_cross(Vec2 a, Vec2 b) { return a.x * b.y - a.y * b.x; }
inline Vec2 b2_cross_vs(Vec2 a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 b2_cross_sv(float s, Vec2 a) { return Vec2(-s * a.y, s * a.x); }
inline Vec2 b2_mul(const Mat22& a, Vec2 v) { return Vec2(a.ex.x * v.x + a.ey.x * v.y, a.ex.y * v.x + a.ey.y * v.y); }
inline float b2_distance_squared(Vec2 a, Vec2 b) { Vec2 c = a - b; return b2_dot(c, c); }
// :612-680
inline Rot b2_mul_t_rot(Rot q, Rot r) { Rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }
inline Vec2 b2_mul_rot(Rot q, Vec2 v) { return Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline Vec2 b2_mul_t_rot(Rot q, Vec2 v) { return Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline Vec2 b2_mul_xf(const Transform& t, Vec2 v) {
  float x = (t.q.c * v.x - t.q.s * v.y) + t.p.x;
  float y = (t.q.s * v.x + t.q.c * v.y) + t.p.y;
  return Vec2(x, y);
}
inline Vec2 b2_mul_t_xf(const Transform& t, Vec2 v) {
  float px = v.x - t.p.x, py = v.y - t.p.y;
  return Vec2(t.q.c * px + t.q.s * py, -t.q.s * px + t.q.c * py);
}
inline Transform b2_mul_t_xf_xf(const Transform& a, const Transform& b) {
  Transform c;
  c.q = b2_mul_t_rot(a.q, b.q);
  c.p = b2_mul_t_rot(a.q, b.p - a.p);
  return c;
}
inline Vec2 b2_min_v(Vec2 a, Vec2 b) { return Vec2(b2_min(a.x, b.x), b2_min(a.y, b.y)); }
inline Vec2 b2_max_v(Vec2 a, Vec2 b) { return Vec2(b2_max(a.x, b.x), b2_max(a.y, b.y)); }

// src/b2_collision.rs:201-256, :355-368
struct AABB {
  Vec2 lower, upper;
  float get_perimeter() const {
    float wx = upper.x - lower.x, wy = upper.y - lower.y;
    return 2.0f * (wx + wy);
  }
  Vec2 get_center() const { return 0.5f * (lower + upper); }
  void combine_two(const AABB& a, const AABB& b) {
    lower = b2_min_v(a.lower, b.lower);
    upper = b2_max_v(a.upper, b.upper);
  }
  bool contains(const AABB& a) const {
    bool r = true;
    r = r && lower.x <= a.lower.x;
    r = r && lower.y <= a.lower.y;
    r = r && a.upper.x <= upper.x;
    r = r && a.upper.y <= upper.y;
    return r;
  }
};
inline bool b2_test_overlap(const AABB& a, const AABB& b) {
  Vec2 d1 = b.lower - a.upper, d2 = a.lower - b.upper;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}

}  // namespace b2o
