// Microbenchmark: does prefetch.global.L1 (no destination register, no scoreboard) make a later plain load an L1 hit?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o prefetch_l1 prefetch_l1.cu && ./prefetch_l1
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE> __global__ void k(const float4* buf, int n_slots, long long* cyc, float* out) {
  unsigned slot = 777u + threadIdx.x;
  float acc = 0.f, spin = 1.0f;
  long long total = 0;
  for (int i = 0; i < ITERS; ++i) {
    slot = slot * 1664525u + 1013904223u;
    const float4* p = buf + ((slot >> 8) % (unsigned)n_slots) * 8;  // a line this SM has not touched (n_slots lines)
    if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (MODE == 3) { float4 v = *p; acc += v.x; }                        // a plain load beforehand: surely in L1 afterwards
    if (MODE == 4) { float4 v = __ldcg(p); acc += v.x; }                 // an L2-only load beforehand
#pragma unroll 1
    for (int j = 0; j < 400; ++j) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(spin) : "f"(1.0000001f));  // ~1600 cycles
    long long t0 = clock64();
    float4 v = *p;
    acc += v.x + v.w;
    asm volatile("" ::"f"(acc));
    long long t1 = clock64();
    total += t1 - t0;
  }
  if (threadIdx.x == 0) { cyc[0] = total; }
  out[threadIdx.x] = acc + spin;
}
template <int MODE> void run(const char* name, const float4* buf, int n_slots, long long* cyc, float* out) {
  k<MODE><<<1, 1>>>(buf, n_slots, cyc, out);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-60s %8.1f cycles from the load to its first use\n", name, (double)h / ITERS);
}
int main() {
  float4* buf; long long* cyc; float* out;
  const int lines = 1 << 18;  // 32 MB: in L2, far more than L1
  cudaMalloc(&buf, (size_t)lines * 128); cudaMemset(buf, 0, (size_t)lines * 128);
  cudaMalloc(&cyc, 8); cudaMalloc(&out, 1024 * 4);
  run<0>("no prefetch (L2 hit)", buf, lines, cyc, out);
  run<1>("prefetch.global.L1 1600 cycles earlier", buf, lines, cyc, out);
  run<2>("prefetch.global.L2 1600 cycles earlier", buf, lines, cyc, out);
  run<3>("plain load 1600 cycles earlier (L1 hit)", buf, lines, cyc, out);
  run<4>("ld.cg 1600 cycles earlier", buf, lines, cyc, out);
  printf("rc=%d\n", (int)cudaGetLastError());
  return 0;
}
