// Microbenchmark: a read-modify-write chain handed from thread to thread of ONE CTA through memory, one barrier per step —
// the pattern of the level-scheduled island sweeps (b2g_levels.h): thread A loads a body, computes, stores it; after the
// barrier thread B loads the same body.  Which load / store flavours keep the hand-over in L1?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o raw_chain raw_chain.cu && ./raw_chain
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
__device__ __forceinline__ float4 ld_plain(const float4* p) { return *p; }
template <int LD, int ST, int SLOTS> __global__ void k(float4* buf, long long* cyc, float* out) {
  __shared__ float4 sm[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 64) sm[tid] = make_float4(0, 0, 0, 0);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
    const int s = (i % SLOTS) * 8;      // SLOTS distinct lines (8 float4 = 128 B apart), revisited round-robin
    const int worker = (i * 37) % nt;   // a different thread (usually another warp) every step
    if (tid == worker) {
      float4 v;
      if (LD == 0) v = buf[s];
      if (LD == 1) v = __ldcg(&buf[s]);
      if (LD == 2) v = sm[s / 8 % 64];
      if (LD == 3) v = __ldcv(&buf[s]);
      v.x += 1.0f; v.y = v.x * 0.5f; v.z += v.y; v.w = (float)i;
      if (ST == 0) buf[s] = v;
      if (ST == 1) __stcg(&buf[s], v);
      if (ST == 2) sm[s / 8 % 64] = v;
      if (ST == 3) __stwt(&buf[s], v);
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) { cyc[0] = t1 - t0; out[0] = buf[0].x + sm[0].x; }
}
template <int LD, int ST, int SLOTS> void run(const char* name, float4* buf, long long* cyc, float* out) {
  cudaMemset(buf, 0, 1 << 20);
  k<LD, ST, SLOTS><<<1, 256>>>(buf, cyc, out);
  cudaMemset(buf, 0, 1 << 20);
  k<LD, ST, SLOTS><<<1, 256>>>(buf, cyc, out);
  cudaDeviceSynchronize();
  long long h; float o;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(&o, out, 4, cudaMemcpyDeviceToHost);
  printf("%-66s %8.1f cycles per step   (check %.0f)\n", name, (double)h / ITERS, o);
}
int main() {
  float4* buf; long long* cyc; float* out;
  cudaMalloc(&buf, 1 << 20); cudaMalloc(&cyc, 8); cudaMalloc(&out, 64);
  run<2, 2, 1>("shared memory, 1 slot", buf, cyc, out);
  run<0, 0, 1>("global: plain ld, plain st, the same slot every step", buf, cyc, out);
  run<0, 0, 8>("global: plain ld, plain st, 8 slots round-robin", buf, cyc, out);
  run<0, 0, 512>("global: plain ld, plain st, 512 slots round-robin", buf, cyc, out);
  run<1, 0, 8>("global: ld.cg, plain st, 8 slots", buf, cyc, out);
  run<0, 1, 8>("global: plain ld, st.cg, 8 slots", buf, cyc, out);
  run<1, 1, 8>("global: ld.cg, st.cg, 8 slots", buf, cyc, out);
  run<3, 3, 8>("global: ld.cv, st.wt, 8 slots", buf, cyc, out);
  run<0, 3, 8>("global: plain ld, st.wt, 8 slots", buf, cyc, out);
  printf("rc=%d\n", (int)cudaGetLastError());
  return 0;
}
