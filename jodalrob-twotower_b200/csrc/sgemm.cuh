// fp32 SIMT GEMM main loop shared by the tower Linear kernels and the fp32 sim-CE kernels.
//
//   acc[m][n] = sum_r A(m,r) * B(n,r)        64x64 tile per 256-thread CTA, 4x4 per thread, BK=16
//
// A(m,r) lives at A[m*lda + r] (R-contiguous, "KC") or A[r*lda + m] (M-contiguous); same for B.
// This is the exact-fp32 path (1e-5 parity vs the CPU oracle); the bf16 path uses tcgen05.
#pragma once
#include "common.cuh"

namespace tt {

constexpr int GM = 64, GN = 64, GK = 16, GPAD = 4, GTHREADS = 256;

struct GemmSmem {
  float a[2][GK][GM + GPAD];
  float b[2][GK][GN + GPAD];
};

// load a 64(rows) x 16(r) operand tile into registers (4 floats per thread)
template <bool KC>
__device__ __forceinline__ void gemm_ldg(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t rows, int64_t r0,
                                         int64_t r_end, bool vec_ok, float (&v)[4], int tid) {
  if (KC) {
    // element (row, r): 4 threads per row, each one float4 along r
    const int row = tid >> 2, c4 = (tid & 3) * 4;
    const int64_t gr = row0 + row, r = r0 + c4;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (gr < rows) {
      const float* p = P + gr * ld + r;
      if (vec_ok && r + 3 < r_end) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (r + j < r_end) v[j] = __ldg(p + j);
      }
    }
  } else {
    // element (r, row): 16 threads per r, each one float4 along rows
    const int rr = tid >> 4, m4 = (tid & 15) * 4;
    const int64_t r = r0 + rr, gr = row0 + m4;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (r < r_end) {
      const float* p = P + r * ld + gr;
      if (vec_ok && gr + 3 < rows) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gr + j < rows) v[j] = __ldg(p + j);
      }
    }
  }
}

template <bool KC, int W>
__device__ __forceinline__ void gemm_sts(float (*s)[W], const float (&v)[4], int tid) {
  if (KC) {
    const int row = tid >> 2, c4 = (tid & 3) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) s[c4 + j][row] = v[j];
  } else {
    const int rr = tid >> 4, m4 = (tid & 15) * 4;
    *reinterpret_cast<float4*>(&s[rr][m4]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// acc += tile product over r in [r_begin, r_end).  Thread (ty, tx) = (tid/16, tid%16) owns rows
// ty*4..+3 and columns tx*4..+3 of the 64x64 tile.
template <bool A_KC, bool B_KC>
__device__ __forceinline__ void gemm_mainloop(const float* __restrict__ A, int64_t lda, int64_t m0, int64_t M,
                                              const float* __restrict__ B, int64_t ldb, int64_t n0, int64_t N,
                                              int64_t r_begin, int64_t r_end, GemmSmem& sm, float (&acc)[4][4]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((r_begin & 3) == 0 || !A_KC);
  const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((r_begin & 3) == 0 || !B_KC);
  float ra[4], rb[4];
  gemm_ldg<A_KC>(A, lda, m0, M, r_begin, r_end, a_vec, ra, tid);
  gemm_ldg<B_KC>(B, ldb, n0, N, r_begin, r_end, b_vec, rb, tid);
  gemm_sts<A_KC>(sm.a[0], ra, tid);
  gemm_sts<B_KC>(sm.b[0], rb, tid);
  __syncthreads();
  int buf = 0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += GK) {
    const bool more = r0 + GK < r_end;
    if (more) {
      gemm_ldg<A_KC>(A, lda, m0, M, r0 + GK, r_end, a_vec, ra, tid);
      gemm_ldg<B_KC>(B, ldb, n0, N, r0 + GK, r_end, b_vec, rb, tid);
    }
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&sm.a[buf][k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sm.b[buf][k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      gemm_sts<A_KC>(sm.a[buf ^ 1], ra, tid);
      gemm_sts<B_KC>(sm.b[buf ^ 1], rb, tid);
    }
    __syncthreads();
    buf ^= 1;
  }
}

// C[m*ldc+n] = act(alpha * acc + bias[n])
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(GTHREADS)
sgemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
             const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t R,
             int64_t r_chunk, int64_t split_stride, float alpha, int relu) {
  __shared__ GemmSmem sm;
  const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  const int64_t r_begin = (int64_t)blockIdx.z * r_chunk;
  const int64_t r_end = (r_begin + r_chunk < R) ? r_begin + r_chunk : R;
  float acc[4][4] = {};
  gemm_mainloop<A_KC, B_KC>(A, lda, m0, M, B, ldb, n0, N, r_begin, r_end, sm, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float* Cz = C + (int64_t)blockIdx.z * split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      Cz[m * ldc + n] = v;
    }
  }
}

}  // namespace tt
