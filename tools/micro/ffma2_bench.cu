// Microbenchmark: latency and single-warp issue rate of scalar fp32 ops against the packed f32x2 forms of
// sm_100a (FFMA2).  One warp per CTA, one CTA per SM -- the occupancy of the batched solver kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

__constant__ float2 k_one = {1.0f, 1.0f};
template <int MODE> __global__ void k(float* out, long long* cyc, float seed) {
  float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  u64 p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7), p4 = pk(a1, a0), p5 = pk(a3, a2), p6 = pk(a5, a4), p7 = pk(a7, a6);
  const float c = 1.0000001f;
  const u64 C = pk(c, c), Z = pk(-0.0f, -0.0f);
  const u64 O = pk(k_one.x, k_one.y);
  const u64 CR = pk(c * seed, c + seed - 1.0f), ZR = pk(seed - 1.0f, 1.0f - seed);  // register operands (seed is a kernel argument)
  int i0 = (int)seed, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
  const int ic = 0x55;
  double d0 = seed, d1 = seed + 1, d2 = seed + 2, d3 = seed + 3, d4 = seed + 4, d5 = seed + 5, d6 = seed + 6, d7 = seed + 7;
  const double dc = 0.999999;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {  // dependent scalar FMUL chain, 8 per iteration
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
    } else if (MODE == 1) {  // dependent FFMA2 chain (mul as fma(a, c, -0))
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(C), "l"(Z));
    } else if (MODE == 2) {  // 8 independent scalar FMUL
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a1) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a2) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a3) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a4) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a5) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a6) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a7) : "f"(c));
    } else if (MODE == 3) {  // 8 independent FFMA2
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p4) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p5) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p6) : "l"(C), "l"(Z));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p7) : "l"(C), "l"(Z));
    } else if (MODE == 4) {  // 8 independent scalar FADD
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a1) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a2) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a3) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a4) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a5) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a6) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a7) : "f"(c));
    } else if (MODE == 5) {  // 4 scalar FMUL + 4 scalar FADD independent, alternating
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a1) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a2) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a3) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a4) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a5) : "f"(c));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a6) : "f"(c));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a7) : "f"(c));
    } else if (MODE == 6) {  // dependent scalar FADD chain
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
    } else if (MODE == 8) {  // 4 FMUL + 4 LOP3 independent, alternating (fma pipe + alu pipe)
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i1) : "r"(ic));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a2) : "f"(c));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i3) : "r"(ic));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a4) : "f"(c));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i5) : "r"(ic));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a6) : "f"(c));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i7) : "r"(ic));
    } else if (MODE == 9) {  // 8 independent LOP3
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i0) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i1) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i2) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i3) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i4) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i5) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i6) : "r"(ic));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(i7) : "r"(ic));
    } else if (MODE == 10) {  // 8 independent DFMA
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d0) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d1) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d2) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d3) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d4) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d5) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d6) : "d"(dc));
      asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d7) : "d"(dc));
    } else if (MODE == 11) {  // dependent DFMA chain
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d0) : "d"(dc));
    } else if (MODE == 12) {  // 4 FMUL + 4 FFMA2 independent, alternating
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(C), "l"(Z));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a2) : "f"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(C), "l"(Z));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a4) : "f"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p5) : "l"(C), "l"(Z));
      asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a6) : "f"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p7) : "l"(C), "l"(Z));
    } else if (MODE == 13) {  // 8 independent FFMA2 with an opaque (constant-bank) multiplier: the form the solver uses for a + b
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p4) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p5) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p6) : "l"(O), "l"(C));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p7) : "l"(O), "l"(C));
    } else if (MODE == 14) {  // 8 independent FMUL2 R, R, R
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p4) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p5) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p6) : "l"(CR));
      asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p7) : "l"(CR));
    } else if (MODE == 15) {  // 8 independent FADD2 R, R, R
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p4) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p5) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p6) : "l"(CR));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p7) : "l"(CR));
    } else if (MODE == 16) {  // 8 independent FFMA2 R, R, R, R
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p4) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p5) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p6) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p7) : "l"(CR), "l"(ZR));
    } else if (MODE == 17) {  // 8 independent FFMA2 R, R, R, R with the accumulator as addend
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p0) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p1) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p2) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p3) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p4) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p5) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p6) : "l"(CR), "l"(ZR));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p7) : "l"(CR), "l"(ZR));
    } else if (MODE == 7) {  // dependent chain alternating scalar FMUL -> FFMA2 consuming it -> scalar of its half
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a0) : "f"(c));
        u64 t = pk(a0, a0);
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(t) : "l"(C), "l"(Z));
        float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(t));
        a0 = x;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 32 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + lo(p0) + lo(p1) + lo(p2) + lo(p3) + lo(p4) + lo(p5) + lo(p6) + lo(p7) + (float)(i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7) + (float)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
}
template <int MODE> void run(const char* name, float* out, long long* cyc, int threads = 32) {
  k<MODE><<<148, threads>>>(out, cyc, 1.0f);
  k<MODE><<<148, threads>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
  printf("%-58s %7.2f cycles per instruction (8 per iteration)\n", name, s / 148 / ITERS / 8);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("dependent FMUL chain (latency)", out, cyc);
  run<6>("dependent FADD chain (latency)", out, cyc);
  run<1>("dependent FFMA2 chain (latency)", out, cyc);
  run<2>("8 independent FMUL, one warp (issue interval)", out, cyc);
  run<4>("8 independent FADD, one warp (issue interval)", out, cyc);
  run<5>("4 FMUL + 4 FADD independent, alternating", out, cyc);
  run<3>("8 independent FFMA2, one warp (issue interval)", out, cyc);
  run<7>("chain FMUL -> FFMA2 -> (4+4 per iteration)", out, cyc);
  run<2>("8 independent FMUL, a 16-thread warp", out, cyc, 16);
  run<3>("8 independent FFMA2, a 16-thread warp", out, cyc, 16);
  run<9>("8 independent LOP3, one warp", out, cyc);
  run<8>("4 FMUL + 4 LOP3 independent, alternating", out, cyc);
  run<12>("4 FMUL + 4 FFMA2 independent, alternating", out, cyc);
  run<13>("8 independent FFMA2, multiplier from the constant bank", out, cyc);
  run<14>("8 independent FMUL2 R,R,R (register operands)", out, cyc);
  run<15>("8 independent FADD2 R,R,R", out, cyc);
  run<16>("8 independent FFMA2 R,R,R,R", out, cyc);
  run<17>("8 independent FFMA2 R,R,R,R (accumulating)", out, cyc);
  run<11>("dependent DFMA chain (latency)", out, cyc);
  run<10>("8 independent DFMA, one warp (issue interval)", out, cyc);
  printf("rc=%d\n", (int)cudaGetLastError());
  return 0;
}
