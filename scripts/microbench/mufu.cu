// Micro-benchmark: ex2.approx.ftz.f32 throughput per SM, alone and mixed with the sim-CE epilogue's FP32 ops.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float c1, float c0) {
  float a[8], acc[4] = {0, 0, 0, 0}, accs[4] = {0, 0, 0, 0}, m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float raw = a[j];
      if (MODE == 0) { acc[j & 3] += ex2f(raw); }                                   // MUFU + FADD
      if (MODE == 1) { acc[j & 3] += ex2f(fmaf(raw, c1, c0)); accs[j & 3] += raw; m4[j & 3] = fmaxf(m4[j & 3], raw); }  // full fwd mix
      if (MODE == 2) { acc[j & 3] += fmaf(raw, c1, c0); accs[j & 3] += raw; m4[j & 3] = fmaxf(m4[j & 3], raw); }         // no MUFU
      a[j] = raw + 1e-6f;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 4; ++j) s += acc[j] + accs[j] + m4[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[100000] = t1 - t0;
}
int main() {
  float* d; cudaMalloc(&d, 1 << 22);
  const int iters = 4000;
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.01f, 0.1f);
      if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, 1.01f, 0.1f);
      if (mode == 2) k<2><<<148, warps * 32>>>(d, iters, 1.01f, 0.1f);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, (char*)d + 100000 * 8, 8, cudaMemcpyDeviceToHost);
      double elems = (double)iters * 8 * warps * 32;
      printf("warps/SM %2d mode %d: %.2f elements/clk/SM  (%.1f clk per warp-element-instr group)\n", warps, mode, elems / c, (double)c / (iters * 8));
    }
  }
  return 0;
}
