// Microbenchmark: latency of a store -> __syncthreads -> load hand-over between two threads of ONE CTA through global
// memory (the level-scheduled island sweeps of b2g_levels.h hand body state over this way), against shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o raw_latency raw_latency.cu && ./raw_latency
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE, int RD> __global__ void k(float4* buf, int n_slots, long long* cyc, float* out) {
  __shared__ float4 sm[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  float4 v = make_float4(tid, 1, 2, 3);
  unsigned slot = 12345u;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    slot = slot * 1664525u + 1013904223u;
    const int s = (int)((slot >> 8) % (unsigned)n_slots);
    const int writer = i % nt, reader = (i + RD) % nt;
    if (tid == writer) {
      v.x += 1.0f;
      if (MODE == 0) buf[s] = v;                                  // plain store
      if (MODE == 1) __stcg(&buf[s], v);                          // st.global.cg
      if (MODE == 2) sm[s & 1023] = v;                            // shared
      if (MODE == 3) buf[s] = v;
      if (MODE == 4) __stwt(&buf[s], v);
    }
    __syncthreads();
    if (tid == reader) {
      float4 r;
      if (MODE == 0) r = buf[s];                                  // plain load
      if (MODE == 1) r = __ldcg(&buf[s]);                         // ld.global.cg
      if (MODE == 2) r = sm[s & 1023];
      if (MODE == 3) { volatile float* p = (volatile float*)&buf[s]; r = make_float4(p[0], p[1], p[2], p[3]); }
      if (MODE == 4) r = __ldcv(&buf[s]);
      acc += r.x + r.w;
      v.y = acc * 1e-30f;  // the next store depends on the load: a chain
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  out[tid] = acc + v.y;
}
template <int MODE, int RD = 7> void run(const char* name, float4* buf, int n_slots, long long* cyc, float* out) {
  k<MODE, RD><<<1, 256>>>(buf, n_slots, cyc, out);
  k<MODE, RD><<<1, 256>>>(buf, n_slots, cyc, out);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-64s %8.1f cycles per hand-over (store, barrier, load, barrier)\n", name, (double)h / ITERS);
}
int main() {
  float4* buf; long long* cyc; float* out;
  const int big = 100000, small = 512;
  cudaMalloc(&buf, big * 16); cudaMemset(buf, 0, big * 16);
  cudaMalloc(&cyc, 8); cudaMalloc(&out, 1024 * 4);
  run<2>("shared memory", buf, small, cyc, out);
  run<0>("global, plain st / ld, 512 slots (8 KB)", buf, small, cyc, out);
  run<0>("global, plain st / ld, 100k slots (1.6 MB)", buf, big, cyc, out);
  run<1>("global, st.cg / ld.cg, 512 slots", buf, small, cyc, out);
  run<1>("global, st.cg / ld.cg, 100k slots", buf, big, cyc, out);
  run<3>("global, plain st / volatile ld, 100k slots", buf, big, cyc, out);
  run<4>("global, st.wt / ld.cv, 100k slots", buf, big, cyc, out);
  run<0, 0>("global plain, reader = writer (same thread), 100k slots", buf, big, cyc, out);
  run<0, 1>("global plain, reader = writer + 1 (same warp mostly), 100k", buf, big, cyc, out);
  run<0, 32>("global plain, reader = writer + 32 (next warp), 100k", buf, big, cyc, out);
  run<0, 96>("global plain, reader = writer + 96 (another scheduler), 100k", buf, big, cyc, out);
  run<2, 32>("shared, reader = writer + 32", buf, small, cyc, out);
  run<1, 32>("global cg, reader = writer + 32, 100k", buf, big, cyc, out);
  printf("rc=%d\n", (int)cudaGetLastError());
  return 0;
}
