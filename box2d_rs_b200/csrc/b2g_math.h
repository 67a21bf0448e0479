// b2g_math.h — fp32 vector algebra of the step path, usable from host and device.
//
// Semantics follow box2d-rs src/b2_math.rs (B2vec2 :23-97, B2Mat22 :261-274, B2Rot :355-376,
// b2_dot/b2_cross/b2_mul* :470-680, b2_min/b2_max/b2_clamp :704-731) and src/b2_common.rs:25-91.
// Every expression keeps the reference's operation order; the translation unit is compiled with
// --fmad=false (nvcc) / -ffp-contract=off (g++) so each fp32 operation rounds exactly once, like
// rustc's output.  min/max are compare-select (NOT fminf/fmaxf: they differ on signed zero / NaN).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B2G_HD __host__ __device__ __forceinline__
#define B2G_HDN __host__ __device__
#else
#define B2G_HD inline
#define B2G_HDN
#endif

namespace b2g {

// src/b2_common.rs:17-91, src/b2_settings.rs:7-11
#define B2G_MAX_FLOAT FLT_MAX
#define B2G_EPSILON FLT_EPSILON
#define B2G_PI 3.14159265358979323846f
#define B2G_AABB_EXTENSION 0.1f
#define B2G_AABB_MULTIPLIER 4.0f
#define B2G_LINEAR_SLOP 0.005f
#define B2G_POLYGON_RADIUS (2.0f * B2G_LINEAR_SLOP)
#define B2G_MAX_LINEAR_CORRECTION 0.2f
#define B2G_MAX_TRANSLATION 2.0f
#define B2G_MAX_TRANSLATION_SQUARED (B2G_MAX_TRANSLATION * B2G_MAX_TRANSLATION)
#define B2G_MAX_ROTATION (0.5f * B2G_PI)
#define B2G_MAX_ROTATION_SQUARED (B2G_MAX_ROTATION * B2G_MAX_ROTATION)
#define B2G_BAUMGARTE 0.2f
#define B2G_TIME_TO_SLEEP 0.5f
#define B2G_LINEAR_SLEEP_TOLERANCE 0.01f
#define B2G_ANGULAR_SLEEP_TOLERANCE (2.0f / 180.0f * B2G_PI)
#define B2G_MAX_POLY 8

B2G_HD float fmin_sel(float a, float b) { return a < b ? a : b; }  // b2_min
// Branch-free select that stays one: the mask is opaque to the optimiser, so (a & m) | (b & ~m) is not
// turned back into a select and from there into a branch (a branch ends a scheduling region on the
// latency-bound chains of the Gauss-Seidel kernels).
B2G_HD int sel_mask(bool c) {
  int m = c ? -1 : 0;
#if defined(__CUDA_ARCH__)
  asm("" : "+r"(m));
#endif
  return m;
}
B2G_HD float msel(int m, float a, float b) {
#if defined(__CUDA_ARCH__)
  return __int_as_float((__float_as_int(a) & m) | (__float_as_int(b) & ~m));
#else
  return m ? a : b;
#endif
}
B2G_HD float fmax_sel(float a, float b) { return a > b ? a : b; }  // b2_max
B2G_HD float fclamp_sel(float a, float lo, float hi) { return fmax_sel(lo, fmin_sel(a, hi)); }
B2G_HD int imin(int a, int b) { return a < b ? a : b; }
B2G_HD int imax(int a, int b) { return a > b ? a : b; }

struct V2 {
  float x, y;
};
B2G_HD V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
B2G_HD V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
B2G_HD V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
B2G_HD V2 operator-(V2 a) { return v2(-a.x, -a.y); }
B2G_HD V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
B2G_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
B2G_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
B2G_HD V2 cross_vs(V2 a, float s) { return v2(s * a.y, -s * a.x); }
B2G_HD V2 cross_sv(float s, V2 a) { return v2(-s * a.y, s * a.x); }
B2G_HD float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
B2G_HD float dist_sq(V2 a, V2 b) { V2 c = a - b; return dot(c, c); }
B2G_HD V2 vmin(V2 a, V2 b) { return v2(fmin_sel(a.x, b.x), fmin_sel(a.y, b.y)); }
B2G_HD V2 vmax(V2 a, V2 b) { return v2(fmax_sel(a.x, b.x), fmax_sel(a.y, b.y)); }
// B2vec2::normalize (src/b2_math.rs:82-92): leaves the vector alone below epsilon, multiplies by 1/len.
B2G_HD float normalize(V2& a) {
  float len = length(a);
  if (len < B2G_EPSILON) return 0.0f;
  float inv = 1.0f / len;
  a.x *= inv;
  a.y *= inv;
  return len;
}

struct Rot {
  float s, c;
};
struct Xf {
  V2 p;
  Rot q;
};
B2G_HD V2 rot_mul(Rot q, V2 v) { return v2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
B2G_HD V2 rot_mul_t(Rot q, V2 v) { return v2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
B2G_HD Rot rot_mul_t_rot(Rot q, Rot r) { Rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }
B2G_HD V2 xf_mul(const Xf& t, V2 v) {
  float x = (t.q.c * v.x - t.q.s * v.y) + t.p.x;
  float y = (t.q.s * v.x + t.q.c * v.y) + t.p.y;
  return v2(x, y);
}
B2G_HD V2 xf_mul_t(const Xf& t, V2 v) {
  float px = v.x - t.p.x, py = v.y - t.p.y;
  return v2(t.q.c * px + t.q.s * py, -t.q.s * px + t.q.c * py);
}
B2G_HD Xf xf_mul_t_xf(const Xf& a, const Xf& b) {
  Xf c;
  c.q = rot_mul_t_rot(a.q, b.q);
  c.p = rot_mul_t(a.q, b.p - a.p);
  return c;
}

struct Box {  // B2AABB, src/b2_collision.rs:201-256
  V2 lo, hi;
};
B2G_HD float box_perimeter(const Box& b) {
  float wx = b.hi.x - b.lo.x, wy = b.hi.y - b.lo.y;
  return 2.0f * (wx + wy);
}
B2G_HD V2 box_center(const Box& b) { return 0.5f * (b.lo + b.hi); }
B2G_HD Box box_union(const Box& a, const Box& b) { Box r; r.lo = vmin(a.lo, b.lo); r.hi = vmax(a.hi, b.hi); return r; }
B2G_HD bool box_contains(const Box& outer, const Box& a) {
  return outer.lo.x <= a.lo.x && outer.lo.y <= a.lo.y && a.hi.x <= outer.hi.x && a.hi.y <= outer.hi.y;
}
// b2_test_overlap (src/b2_collision.rs:355-368): touching boxes overlap.
B2G_HD bool box_overlap(const Box& a, const Box& b) {
  float d1x = b.lo.x - a.hi.x, d1y = b.lo.y - a.hi.y;
  float d2x = a.lo.x - b.hi.x, d2y = a.lo.y - b.hi.y;
  if (d1x > 0.0f || d1y > 0.0f) return false;
  if (d2x > 0.0f || d2y > 0.0f) return false;
  return true;
}

// ------------------------------------------------------------------------------------------
// sinf/cosf.  B2Rot::set (src/b2_math.rs:372-376) calls f32::sin / f32::cos, which lower to the
// platform libm.  On x86-64 linux-gnu that is glibc's sincosf family (sysdeps/ieee754/flt-32,
// the FMA ifunc variant on every CPU of the last decade): a double-precision polynomial after a
// double-precision reduction by pi/2.  The functions below restate that published algorithm
// operation for operation (explicit fma where the FMA build contracts), so the results are
// bit-identical to glibc 2.39's sinf/cosf for all 2^32 inputs — checked exhaustively on the CPU
// by oracle/sincosf_check.c and sampled on the GPU by tests/test_gpu_math.py.
// ------------------------------------------------------------------------------------------
struct SinCosTab {
  double c0, c1, c2, c3, c4, s1, s2, s3;
};
B2G_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
B2G_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ffu; }

// Evaluates sin (odd==0) or cos (odd==1) of the reduced argument; neg selects the table whose
// cosine coefficients are negated ("-cos for free" in quadrants 2,3).
B2G_HD float sc_poly(double x, double x2, int neg, int odd) {
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  if ((odd & 1) == 0) {
    double x3 = x * x2;
    double s1 = fma(x2, S3, S2);
    double x7 = x3 * x2;
    double s = fma(x3, S1, x);
    return (float)fma(x7, s1, s);
  } else {
    double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
           C4 = 0x1.99343027bf8c3p-16;
    if (neg) { C0 = -C0; C1 = -C1; C2 = -C2; C3 = -C3; C4 = -C4; }
    double x4 = x2 * x2;
    double c2 = fma(x2, C4, C3);
    double c1 = fma(x2, C1, C0);
    double x6 = x4 * x2;
    double c = fma(x4, C2, c1);
    return (float)fma(x6, c2, c);
  }
}
B2G_HDN inline double sc_reduce_large(uint32_t xi, int* np) {
  const uint32_t inv_pio4[24] = {0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44,
                                 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
                                 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62,
                                 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
  const uint32_t* arr = &inv_pio4[(xi >> 26) & 15];
  int shift = (xi >> 23) & 7;
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  res0 = (uint32_t)(xi * arr[0]);
  res1 = (uint64_t)xi * arr[4];
  res2 = (uint64_t)xi * arr[8];
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  double x = (double)(int64_t)res0;
  *np = (int)n;
  return x * 0x1.921FB54442D18p-62;
}
// sin and cos of one angle (B2Rot::set).  Shares the reduction between the two results, which
// glibc's separate sinf/cosf calls repeat identically.
B2G_HD void sincos_ref(float y, float* sp, float* cp) {
  double x = (double)y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) { *sp = y; *cp = 1.0f; return; }
    *sp = sc_poly(x, x2, 0, 0);
    *cp = sc_poly(x, x2, 0, 1);
    return;
  }
  int n, q;
  if (abstop12(y) < abstop12(120.0f)) {
    double r = x * 0x1.45F306DC9C883p+23;
    n = ((int32_t)r + 0x800000) >> 24;
    x = fma(-(double)n, 0x1.921FB54442D18p0, x);
    q = n;
  } else if (abstop12(y) < abstop12(INFINITY)) {
    uint32_t xi = f2u(y);
    x = sc_reduce_large(xi, &n);
    q = n + (int)(xi >> 31);
  } else {
    *sp = y - y;
    *cp = y - y;
    return;
  }
  double sgn = ((q & 3) == 1 || (q & 3) == 2) ? -1.0 : 1.0;
  int neg = (q & 2) ? 1 : 0;
  double xs = x * sgn, x2 = x * x;
  *sp = sc_poly(xs, x2, neg, n);
  *cp = sc_poly(xs, x2, neg, n ^ 1);
}
B2G_HD Rot rot_from_angle(float a) { Rot q; sincos_ref(a, &q.s, &q.c); return q; }

// Branch-free form of sincos_ref for |y| < 120 (abstop12(y) < abstop12(120.0f)): the medium-range reduction
// with n = 0 is the identity, so it also serves the small range; both polynomials are evaluated once and
// swapped / negated by selects.  Bit-identical to sincos_ref on its whole domain: tests/test_sincos_mid.py
// compares all 2.2e9 floats of the domain (tools/sincos_mid_check.cpp).
B2G_HD bool sincos_mid_domain(float y) { return abstop12(y) < abstop12(120.0f); }
B2G_HD void sincos_mid(float y, float* sp, float* cp) {
  const double x0 = (double)y;
  const double r = x0 * 0x1.45F306DC9C883p+23;
  const int n = ((int32_t)r + 0x800000) >> 24;
  const double x = fma(-(double)n, 0x1.921FB54442D18p0, x0);
  const bool flip = ((n & 3) == 1 || (n & 3) == 2);
  const double xs = flip ? -x : x;  // x * -1.0 is exact
  const double x2 = x * x;
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
               C4 = 0x1.99343027bf8c3p-16;
  const double x3 = xs * x2;
  const double s1 = fma(x2, S3, S2);
  const double x7 = x3 * x2;
  const double sa = fma(x3, S1, xs);
  const float sn = (float)fma(x7, s1, sa);
  const double x4 = x2 * x2;
  const double c2 = fma(x2, C4, C3);
  const double c1 = fma(x2, C1, C0);
  const double x6 = x4 * x2;
  const double ca = fma(x4, C2, c1);
  const float cpos = (float)fma(x6, c2, ca);
  const float cs = (n & 2) ? -cpos : cpos;  // negated coefficients give the exactly negated value
  const bool odd = (n & 1) != 0;
  const bool tiny = abstop12(y) < abstop12(0x1p-12f);
  const float so = odd ? cs : sn, co = odd ? sn : cs;
  *sp = tiny ? y : so;
  *cp = tiny ? 1.0f : co;
}

}  // namespace b2g
