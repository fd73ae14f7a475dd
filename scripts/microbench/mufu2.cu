// Micro-benchmark (round 2): what bounds the sim-CE epilogues once the per-logit FFMA is gone — MUFU.EX2 alone, with the bf16 pack
// of the backward (F2FP vs integer pack), packed bf16x2 exp2, and the FMA-pipe polynomial mixed in. Elements per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t prmt_hi(uint32_t a, uint32_t b) { uint32_t y; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(y) : "r"(a), "r"(b)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.009574027f, f, 0.055929556f);
  p = fmaf(p, f, 0.24024744f);
  p = fmaf(p, f, 0.69312006f);
  p = fmaf(p, f, 0.99999923f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8], acc[4] = {0, 0, 0, 0};
  uint32_t x = 0;
  for (int j = 0; j < 8; ++j) a[j] = -(threadIdx.x * 1e-3f + j * 0.1f);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float r0 = a[j], r1 = a[j + 1];
      if (MODE == 0) { acc[j & 3] += ex2f(r0); acc[(j + 1) & 3] += ex2f(r1); }
      if (MODE == 1) { x ^= cvt_bf16x2(ex2f(r0), ex2f(r1)); }
      if (MODE == 2) { x ^= prmt_hi(__float_as_uint(ex2f(r0)), __float_as_uint(ex2f(r1))); }
      if (MODE == 3) { x ^= prmt_hi(__float_as_uint(ex2f(r0)) + 0x8000u, __float_as_uint(ex2f(r1)) + 0x8000u); }
      if (MODE == 4) { x ^= ex2_bf16x2(cvt_bf16x2(r0, r1)); }
      if (MODE == 5) { acc[j & 3] += ex2_poly(r0); acc[(j + 1) & 3] += ex2_poly(r1); }
      if (MODE == 6) { acc[j & 3] += ((j & 6) == 6) ? ex2_poly(r0) : ex2f(r0); acc[(j + 1) & 3] += ((j & 6) == 6) ? ex2_poly(r1) : ex2f(r1); }  // 1/4 poly
      if (MODE == 7) { const float e0 = ex2f(r0), e1 = ex2f(r1); acc[j & 3] += e0; acc[(j + 1) & 3] += e1; x ^= prmt_hi(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u); }
      if (MODE == 8) { const float e0 = ex2f(r0), e1 = ex2f(r1); acc[j & 3] += e0; acc[(j + 1) & 3] += e1; x ^= cvt_bf16x2(e0, e1); }
      if (MODE == 9) { const uint32_t e = ex2_bf16x2(cvt_bf16x2(r0, r1)); x ^= e; acc[j & 3] += __uint_as_float(e << 16); acc[(j + 1) & 3] += __uint_as_float(e & 0xffff0000u); }
      a[j] = r0 + 1e-6f; a[j + 1] = r1 + 1e-6f;
    }
  }
  long long t1 = clock64();
  float s = __uint_as_float(x); for (int j = 0; j < 4; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[100000] = t1 - t0;
}
template <int MODE> void run(float* d, int warps, const char* name) {
  const int iters = 4000;
  k<MODE><<<148, warps * 32>>>(d, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, (char*)d + 100000 * 8, 8, cudaMemcpyDeviceToHost);
  printf("warps/SM %2d  %-58s %6.2f exps/clk/SM\n", warps, name, (double)iters * 8 * warps * 32 / c);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 22);
  for (int warps : {16, 32}) {
    run<0>(d, warps, "ex2.f32 + FADD");
    run<1>(d, warps, "2 ex2.f32 + cvt.rn.bf16x2.f32 (F2FP)");
    run<2>(d, warps, "2 ex2.f32 + PRMT (truncating pack)");
    run<3>(d, warps, "2 ex2.f32 + 2 IADD + PRMT (round-half-up pack)");
    run<4>(d, warps, "cvt.rn.bf16x2.f32 + ex2.bf16x2");
    run<5>(d, warps, "polynomial only (FMA pipe)");
    run<6>(d, warps, "3/4 ex2.f32 + 1/4 polynomial, + FADD");
    run<7>(d, warps, "2 ex2.f32 + 2 FADD + 2 IADD + PRMT (P2 with integer pack)");
    run<8>(d, warps, "2 ex2.f32 + 2 FADD + F2FP (P2 now)");
    run<9>(d, warps, "cvt + ex2.bf16x2 + 2 unpack + 2 FADD (P2 packed)");
  }
  return 0;
}
