// Tower Linear layers, exact-fp32 path: forward (+bias, +ReLU), dgrad, split-K wgrad + bias grad.
#include "sgemm.cuh"

namespace tt {

// out[i] = sum_z part[z*stride + i], fixed order
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int64_t stride, int splits, float* __restrict__ out,
                                     int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * stride + i];
  out[i] = s;
}

// column sums of X[M,N] in two fixed levels: chunk partials (float) then a sequential double sum
constexpr int CS_ROWS = 64;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ X, int64_t ld, int64_t M, int N, double* __restrict__ part) {
  __shared__ double s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int64_t m_lo = (int64_t)blockIdx.y * CS_ROWS;
  const int64_t m_hi = (m_lo + CS_ROWS < M) ? m_lo + CS_ROWS : M;
  double a = 0.0;  // bias gradients cancel heavily; keep the column sums in fp64
  if (n < N)
    for (int64_t m = m_lo + ty; m < m_hi; m += 8) a += (double)X[m * ld + n];
  s[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && n < N) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][tx];
    part[(int64_t)blockIdx.y * N + n] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const double* __restrict__ part, int chunks, int N, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per column, fixed shuffle tree
  if (n >= N) return;
  double t = 0.0;
  for (int c = lane; c < chunks; c += 32) t += part[(int64_t)c * N + n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) out[n] = (float)t;
}

static int pick_splits(int64_t M, int tiles) {
  // enough CTAs to cover the machine twice, each split at least 256 rows deep
  int want = (2 * sm_count() + tiles - 1) / tiles;
  int maxs = (int)((M + 255) / 256);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return want;
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_linear_forward(const void* A, int64_t lda, const float* W, const float* bias, void* C, int64_t ldc, int64_t M,
                      int32_t N, int32_t K, int32_t act_dtype, int32_t relu, tt_stream_t stream) {
  TT_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0, "tt_linear_forward: bad arguments");
  TT_CHECK_ARG(lda >= K && ldc >= N, "tt_linear_forward: leading dimension too small");
  if (act_dtype != TT_F32) {
    set_error("tt_linear_forward: act_dtype %d not supported by the fp32 path", act_dtype);
    return TT_ERR_UNSUPPORTED;
  }
  if (M == 0) return TT_OK;
  dim3 grid(ceil_div(N, GN), ceil_div(M, GM), 1);
  sgemm_kernel<true, true><<<grid, GTHREADS, 0, (cudaStream_t)stream>>>((const float*)A, lda, W, K, bias, (float*)C, ldc,
                                                                        M, N, K, K, 0, 1.0f, relu);
  TT_LAUNCH_CHECK("sgemm_kernel<fwd>");
  return TT_OK;
}

int tt_linear_dgrad(const void* dY, int64_t ldy, const float* W, void* dX, int64_t ldx, int64_t M, int32_t N,
                    int32_t K, int32_t act_dtype, tt_stream_t stream) {
  TT_CHECK_ARG(dY && W && dX && M >= 0 && N > 0 && K > 0, "tt_linear_dgrad: bad arguments");
  TT_CHECK_ARG(ldy >= N && ldx >= K, "tt_linear_dgrad: leading dimension too small");
  if (act_dtype != TT_F32) {
    set_error("tt_linear_dgrad: act_dtype %d not supported by the fp32 path", act_dtype);
    return TT_ERR_UNSUPPORTED;
  }
  if (M == 0) return TT_OK;
  // dX[m,k] = sum_n dY[m,n] * W[n,k] : A = dY (reduction contiguous), B(k, n) = W[n*K + k] (row-contiguous)
  dim3 grid(ceil_div(K, GN), ceil_div(M, GM), 1);
  sgemm_kernel<true, false><<<grid, GTHREADS, 0, (cudaStream_t)stream>>>((const float*)dY, ldy, W, K, nullptr,
                                                                         (float*)dX, ldx, M, K, N, N, 0, 1.0f, 0);
  TT_LAUNCH_CHECK("sgemm_kernel<dgrad>");
  return TT_OK;
}

static int wgrad_splits(int64_t M, int32_t N, int32_t K) {
  return pick_splits(M, ceil_div(K, GN) * ceil_div(N, GM));
}

int64_t tt_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K) {
  const int64_t chunks = (M + CS_ROWS - 1) / CS_ROWS + 1;
  return (int64_t)wgrad_splits(M, N, K) * N * K * 4 + chunks * N * 8 + 256;
}

int tt_linear_wgrad(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, float* dbias, int64_t M,
                    int32_t N, int32_t K, int32_t act_dtype, void* workspace, int64_t workspace_bytes,
                    tt_stream_t stream) {
  TT_CHECK_ARG(dY && X && dW && M > 0 && N > 0 && K > 0, "tt_linear_wgrad: bad arguments");
  TT_CHECK_ARG(ldy >= N && ldx >= K, "tt_linear_wgrad: leading dimension too small");
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_linear_wgrad_workspace_bytes(M, N, K), "tt_linear_wgrad: workspace too small");
  if (act_dtype != TT_F32) {
    set_error("tt_linear_wgrad: act_dtype %d not supported by the fp32 path", act_dtype);
    return TT_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  // dW[n,k] = sum_m dY[m,n] * X[m,k] : both operands row(=output index)-contiguous, reduction over m
  const int splits = wgrad_splits(M, N, K);
  int64_t r_chunk = (M + splits - 1) / splits;
  r_chunk = (r_chunk + GK - 1) / GK * GK;
  const int eff_splits = (int)((M + r_chunk - 1) / r_chunk);
  float* part = (float*)workspace;
  dim3 grid(ceil_div(K, GN), ceil_div(N, GM), eff_splits);
  if (eff_splits == 1) {
    sgemm_kernel<false, false><<<grid, GTHREADS, 0, s>>>((const float*)dY, ldy, (const float*)X, ldx, nullptr, dW, K, N,
                                                         K, M, r_chunk, 0, 1.0f, 0);
  } else {
    sgemm_kernel<false, false><<<grid, GTHREADS, 0, s>>>((const float*)dY, ldy, (const float*)X, ldx, nullptr, part, K,
                                                         N, K, M, r_chunk, (int64_t)N * K, 1.0f, 0);
    splitk_reduce_kernel<<<ceil_div((int64_t)N * K, 256), 256, 0, s>>>(part, (int64_t)N * K, eff_splits, dW, (int64_t)N * K);
  }
  if (dbias) {
    double* cpart = reinterpret_cast<double*>(part + (((int64_t)splits * N * K + 1) & ~(int64_t)1));
    const int chunks = (int)((M + CS_ROWS - 1) / CS_ROWS);
    colsum_partial_kernel<<<dim3(ceil_div(N, 32), chunks), 256, 0, s>>>((const float*)dY, ldy, M, N, cpart);
    colsum_final_kernel<<<ceil_div(N, 8), 256, 0, s>>>(cpart, chunks, N, dbias);
  }
  TT_LAUNCH_CHECK("linear_wgrad");
  return TT_OK;
}

}  // extern "C"
