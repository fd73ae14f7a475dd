// BatchNorm1d (train statistics, SyncBN-ready) + Dropout, their backward, and L2 normalisation.
// All are [M, N<=~1k] element-wise / column-reduction passes: bandwidth-trivial, latency-bound, written
// for fixed (deterministic) reduction order.
#include "common.cuh"

namespace tt {

constexpr int BN_ROWS = 64;  // rows per partial chunk (8 rows per thread): enough CTAs to cover the machine

template <typename T>
__device__ __forceinline__ float keep_scale(const uint8_t* mask_in, const tt_adam_hyper* hyp, uint32_t stream,
                                            int64_t m, int n, int N, float p, float inv_keep) {
  if (p <= 0.f) return 1.f;
  const int64_t idx = m * N + n;
  bool keep;
  if (mask_in) keep = mask_in[idx] != 0;
  else keep = dropout_keep(hyp->rng_seed, hyp->rng_step, stream, (uint64_t)idx, p);
  return keep ? inv_keep : 0.f;
}

// MODE 0: (x, x^2)          MODE 1: (dy', dy' * xhat)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
bn_partial_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ H, int64_t ldh, int64_t M, int N,
                  const float* __restrict__ mean, const float* __restrict__ rstd, float p, const uint8_t* mask_in,
                  const tt_adam_hyper* __restrict__ hyp, uint32_t stream, double* __restrict__ part) {
  __shared__ double s0[8][33], s1[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int64_t m_lo = (int64_t)blockIdx.y * BN_ROWS;
  const int64_t m_hi = (m_lo + BN_ROWS < M) ? m_lo + BN_ROWS : M;
  double a = 0.0, b = 0.0;  // fp64 accumulation: sum_b dy cancels heavily and Adam's m/sqrt(v) amplifies the noise
  if (n < N) {
    const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
    const float mu = MODE == 1 ? mean[n] : 0.f, rs = MODE == 1 ? rstd[n] : 0.f;
    for (int64_t m = m_lo + ty; m < m_hi; m += 8) {
      const float x = ldf<T>(X + m * ldx + n);
      if (MODE == 0) {
        a += (double)x;
        b += (double)x * (double)x;
      } else {
        const float dyp = x * keep_scale<T>(mask_in, hyp, stream, m, n, N, p, inv_keep);
        const float xhat = (ldf<T>(H + m * ldh + n) - mu) * rs;
        a += (double)dyp;
        b += (double)dyp * (double)xhat;
      }
    }
  }
  s0[ty][tx] = a;
  s1[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && n < N) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      t0 += s0[k][tx];
      t1 += s1[k][tx];
    }
    part[((int64_t)blockIdx.y * 2 + 0) * N + n] = t0;
    part[((int64_t)blockIdx.y * 2 + 1) * N + n] = t1;
  }
}

// one warp per output column: lanes stride over the chunk partials, then a fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(256) bn_partial_final_kernel(const double* __restrict__ part, int chunks, int N, double* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * 8 + (threadIdx.x >> 5);  // column index in [0, 2N): (which sum) * N + n
  if (col >= 2 * N) return;
  double t = 0.0;
  for (int c = lane; c < chunks; c += 32) t += part[(int64_t)c * 2 * N + col];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) sums[col] = t;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long count, int N, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ rm,
                                   float* __restrict__ rv, long long* __restrict__ nbt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0 && nbt) *nbt += 1;
  if (n >= N) return;
  const double mu = sums[n] / (double)count;
  double var = sums[N + n] / (double)count - mu * mu;  // biased
  if (var < 0.0) var = 0.0;
  mean[n] = (float)mu;
  rstd[n] = (float)(1.0 / sqrt(var + (double)eps));
  if (rm) rm[n] = (1.f - momentum) * rm[n] + momentum * (float)mu;
  if (rv) {
    const double unb = count > 1 ? var * ((double)count / (double)(count - 1)) : var;
    rv[n] = (1.f - momentum) * rv[n] + momentum * (float)unb;
  }
}

__global__ void bn_rstd_kernel(const float* __restrict__ var, int N, float eps, float* __restrict__ rstd) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) rstd[n] = 1.f / sqrtf(var[n] + eps);
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ H, int64_t ldh, int64_t M, int N, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta, float p,
                const uint8_t* mask_in, const tt_adam_hyper* __restrict__ hyp, uint32_t stream, T* __restrict__ Y,
                int64_t ldy) {
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const int64_t total = M * N, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / N;
    const int n = (int)(i - m * N);
    const float xhat = (ldf<T>(H + m * ldh + n) - mean[n]) * rstd[n];
    float y = xhat * gamma[n] + beta[n];
    y *= keep_scale<T>(mask_in, hyp, stream, m, n, N, p, inv_keep);
    stf<T>(Y + m * ldy + n, y);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ DY, int64_t lddy, const T* __restrict__ H, int64_t ldh, int64_t M, int N,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    float p, const uint8_t* mask_in, const tt_adam_hyper* __restrict__ hyp, uint32_t stream,
                    const double* __restrict__ sums, long long count, int train, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float affine_scale, T* __restrict__ DH, int64_t lddh) {
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const int64_t total = M * N, stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_count = 1.f / (float)count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / N;
    const int n = (int)(i - m * N);
    const float h = ldf<T>(H + m * ldh + n);
    const float dyp = ldf<T>(DY + m * lddy + n) * keep_scale<T>(mask_in, hyp, stream, m, n, N, p, inv_keep);
    float dh;
    if (train) {
      const float xhat = (h - mean[n]) * rstd[n];
      const float s0 = (float)sums[n], s1 = (float)sums[N + n];
      dh = rstd[n] * gamma[n] * (dyp - s0 * inv_count - xhat * s1 * inv_count);
    } else {
      dh = rstd[n] * gamma[n] * dyp;
    }
    stf<T>(DH + m * lddh + n, h > 0.f ? dh : 0.f);
    if (i < N) {  // first row's threads also publish the affine grads
      if (dbeta) dbeta[n] = affine_scale * (float)sums[n];
      if (dgamma) dgamma[n] = affine_scale * (float)sums[N + n];
    }
  }
}

// one warp per row
template <typename T>
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ Z, int64_t M, int F, T* __restrict__ E, float* __restrict__ nrm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* z = Z + row * F;
  float ss = 0.f;
  for (int f = lane; f < F; f += 32) ss += z[f] * z[f];
  ss = warp_sum(ss);
  const float n = fmaxf(sqrtf(ss), 1e-12f);
  for (int f = lane; f < F; f += 32) stf<T>(E + row * F + f, z[f] / n);
  if (lane == 0 && nrm) nrm[row] = n;
}

// bf16 tower outputs twice: unit rows, and the same rows times `scale` (rounded once, from the fp32 value): the pre-scaled
// operand of the logits passes (scale = log2(e) / T makes the tensor-core accumulator the exponent of exp2)
__global__ void __launch_bounds__(256)
l2norm_fwd_scaled_kernel(const float* __restrict__ Z, int64_t M, int F, __nv_bfloat16* __restrict__ E, __nv_bfloat16* __restrict__ Es,
                         float scale, float* __restrict__ nrm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* z = Z + row * F;
  float ss = 0.f;
  for (int f = lane; f < F; f += 32) ss += z[f] * z[f];
  ss = warp_sum(ss);
  const float n = fmaxf(sqrtf(ss), 1e-12f);
  for (int f = lane; f < F; f += 32) {
    const float u = z[f] / n;
    E[row * F + f] = __float2bfloat16_rn(u);
    Es[row * F + f] = __float2bfloat16_rn(u * scale);
  }
  if (lane == 0 && nrm) nrm[row] = n;
}

template <typename T>
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ DE, const T* __restrict__ E, const float* __restrict__ nrm, int64_t M, int F,
                  float* __restrict__ DZ) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float dot = 0.f;
  for (int f = lane; f < F; f += 32) dot += ldf<T>(E + row * F + f) * DE[row * F + f];
  dot = warp_sum(dot);
  const float inv = 1.f / nrm[row];
  for (int f = lane; f < F; f += 32) DZ[row * F + f] = (DE[row * F + f] - ldf<T>(E + row * F + f) * dot) * inv;
}

static int ew_grid(int64_t total) {
  int64_t want = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace tt

using namespace tt;

extern "C" {

int64_t tt_bn_workspace_bytes(int64_t M, int32_t N) {
  const int64_t chunks = (M + BN_ROWS - 1) / BN_ROWS + 1;
  return chunks * 2 * (int64_t)N * 8 + 256;
}

int tt_bn_stats(const void* h, int64_t ldh, int64_t M, int32_t N, int32_t act_dtype, double* sums, void* workspace,
                int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(h && sums && workspace && M > 0 && N > 0 && ldh >= N, "tt_bn_stats: bad arguments");
  TT_CHECK_ARG(workspace_bytes >= tt_bn_workspace_bytes(M, N), "tt_bn_stats: workspace too small");
  TT_CHECK_ARG(act_dtype == TT_F32 || act_dtype == TT_BF16, "tt_bn_stats: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  const int chunks = (int)((M + BN_ROWS - 1) / BN_ROWS);
  dim3 grid(ceil_div(N, 32), chunks);
  if (act_dtype == TT_F32)
    bn_partial_kernel<float, 0><<<grid, 256, 0, s>>>((const float*)h, ldh, nullptr, 0, M, N, nullptr, nullptr, 0.f,
                                                     nullptr, nullptr, 0, (double*)workspace);
  else
    bn_partial_kernel<__nv_bfloat16, 0><<<grid, 256, 0, s>>>((const __nv_bfloat16*)h, ldh, nullptr, 0, M, N, nullptr,
                                                             nullptr, 0.f, nullptr, nullptr, 0, (double*)workspace);
  bn_partial_final_kernel<<<ceil_div(2 * N, 8), 256, 0, s>>>((const double*)workspace, chunks, N, sums);
  TT_LAUNCH_CHECK("bn_stats");
  return TT_OK;
}

int tt_bn_finalize(const double* sums, int64_t count, int32_t N, float eps, float momentum, float* mean, float* rstd,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, tt_stream_t stream) {
  TT_CHECK_ARG(sums && mean && rstd && count > 0 && N > 0, "tt_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(sums, (long long)count, N, eps, momentum, mean,
                                                                        rstd, running_mean, running_var,
                                                                        (long long*)num_batches_tracked);
  TT_LAUNCH_CHECK("bn_finalize_kernel");
  return TT_OK;
}

int tt_bn_rstd_from_var(const float* var, int32_t N, float eps, float* rstd, tt_stream_t stream) {
  TT_CHECK_ARG(var && rstd && N > 0, "tt_bn_rstd_from_var: bad arguments");
  bn_rstd_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(var, N, eps, rstd);
  TT_LAUNCH_CHECK("bn_rstd_kernel");
  return TT_OK;
}

int tt_bn_apply(const void* h, int64_t ldh, int64_t M, int32_t N, int32_t act_dtype, const float* mean,
                const float* rstd, const float* gamma, const float* beta, float dropout_p, const uint8_t* mask_in,
                const tt_adam_hyper* d_hyper, uint32_t rng_stream, void* y, int64_t ldy, tt_stream_t stream) {
  TT_CHECK_ARG(h && y && mean && rstd && gamma && beta && M > 0 && N > 0 && ldh >= N && ldy >= N, "tt_bn_apply: bad arguments");
  TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "tt_bn_apply: dropout_p must be in [0,1)");
  TT_CHECK_ARG(dropout_p == 0.f || mask_in || d_hyper, "tt_bn_apply: dropout needs mask_in or d_hyper");
  TT_CHECK_ARG(act_dtype == TT_F32 || act_dtype == TT_BF16, "tt_bn_apply: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (act_dtype == TT_F32)
    bn_apply_kernel<float><<<ew_grid(M * N), 256, 0, s>>>((const float*)h, ldh, M, N, mean, rstd, gamma, beta, dropout_p,
                                                          mask_in, d_hyper, rng_stream, (float*)y, ldy);
  else
    bn_apply_kernel<__nv_bfloat16><<<ew_grid(M * N), 256, 0, s>>>((const __nv_bfloat16*)h, ldh, M, N, mean, rstd, gamma,
                                                                  beta, dropout_p, mask_in, d_hyper, rng_stream,
                                                                  (__nv_bfloat16*)y, ldy);
  TT_LAUNCH_CHECK("bn_apply_kernel");
  return TT_OK;
}

int tt_bn_backward_reduce(const void* dy, int64_t lddy, const void* h, int64_t ldh, int64_t M, int32_t N,
                          int32_t act_dtype, const float* mean, const float* rstd, float dropout_p,
                          const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream, double* sums,
                          void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(dy && h && mean && rstd && sums && workspace && M > 0 && N > 0, "tt_bn_backward_reduce: bad arguments");
  TT_CHECK_ARG(workspace_bytes >= tt_bn_workspace_bytes(M, N), "tt_bn_backward_reduce: workspace too small");
  TT_CHECK_ARG(dropout_p == 0.f || mask_in || d_hyper, "tt_bn_backward_reduce: dropout needs mask_in or d_hyper");
  TT_CHECK_ARG(act_dtype == TT_F32 || act_dtype == TT_BF16, "tt_bn_backward_reduce: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  const int chunks = (int)((M + BN_ROWS - 1) / BN_ROWS);
  dim3 grid(ceil_div(N, 32), chunks);
  if (act_dtype == TT_F32)
    bn_partial_kernel<float, 1><<<grid, 256, 0, s>>>((const float*)dy, lddy, (const float*)h, ldh, M, N, mean, rstd,
                                                     dropout_p, mask_in, d_hyper, rng_stream, (double*)workspace);
  else
    bn_partial_kernel<__nv_bfloat16, 1><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)h, ldh,
                                                             M, N, mean, rstd, dropout_p, mask_in, d_hyper, rng_stream,
                                                             (double*)workspace);
  bn_partial_final_kernel<<<ceil_div(2 * N, 8), 256, 0, s>>>((const double*)workspace, chunks, N, sums);
  TT_LAUNCH_CHECK("bn_backward_reduce");
  return TT_OK;
}

int tt_bn_backward_apply(const void* dy, int64_t lddy, const void* h, int64_t ldh, int64_t M, int32_t N,
                         int32_t act_dtype, const float* mean, const float* rstd, const float* gamma,
                         float dropout_p, const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream,
                         const double* sums, int64_t count, int32_t train, float* dgamma, float* dbeta,
                         float affine_grad_scale, void* dh, int64_t lddh, tt_stream_t stream) {
  TT_CHECK_ARG(dy && h && mean && rstd && gamma && sums && dh && M > 0 && N > 0 && count > 0, "tt_bn_backward_apply: bad arguments");
  TT_CHECK_ARG(dropout_p == 0.f || mask_in || d_hyper, "tt_bn_backward_apply: dropout needs mask_in or d_hyper");
  TT_CHECK_ARG(act_dtype == TT_F32 || act_dtype == TT_BF16, "tt_bn_backward_apply: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (act_dtype == TT_F32)
    bn_bwd_apply_kernel<float><<<ew_grid(M * N), 256, 0, s>>>((const float*)dy, lddy, (const float*)h, ldh, M, N, mean,
                                                              rstd, gamma, dropout_p, mask_in, d_hyper, rng_stream, sums,
                                                              (long long)count, train, dgamma, dbeta, affine_grad_scale, (float*)dh, lddh);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<ew_grid(M * N), 256, 0, s>>>(
        (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)h, ldh, M, N, mean, rstd, gamma, dropout_p, mask_in,
        d_hyper, rng_stream, sums, (long long)count, train, dgamma, dbeta, affine_grad_scale, (__nv_bfloat16*)dh, lddh);
  TT_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return TT_OK;
}

int tt_l2norm_forward(const float* z, int64_t M, int32_t F, void* e, int32_t out_dtype, float* nrm,
                      tt_stream_t stream) {
  TT_CHECK_ARG(z && e && M > 0 && F > 0, "tt_l2norm_forward: bad arguments");
  TT_CHECK_ARG(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_l2norm_forward: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == TT_F32)
    l2norm_fwd_kernel<float><<<ceil_div(M, 8), 256, 0, s>>>(z, M, F, (float*)e, nrm);
  else
    l2norm_fwd_kernel<__nv_bfloat16><<<ceil_div(M, 8), 256, 0, s>>>(z, M, F, (__nv_bfloat16*)e, nrm);
  TT_LAUNCH_CHECK("l2norm_fwd_kernel");
  return TT_OK;
}

int tt_l2norm_forward_scaled(const float* z, int64_t M, int32_t F, void* e_bf16, void* e_scaled_bf16, float scale, float* nrm,
                             tt_stream_t stream) {
  TT_CHECK_ARG(z && e_bf16 && e_scaled_bf16 && M > 0 && F > 0, "tt_l2norm_forward_scaled: bad arguments");
  l2norm_fwd_scaled_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(z, M, F, (__nv_bfloat16*)e_bf16,
                                                                            (__nv_bfloat16*)e_scaled_bf16, scale, nrm);
  TT_LAUNCH_CHECK("l2norm_fwd_scaled_kernel");
  return TT_OK;
}

int tt_l2norm_backward(const float* de, const void* e, int32_t e_dtype, const float* nrm, int64_t M, int32_t F,
                       float* dz, tt_stream_t stream) {
  TT_CHECK_ARG(de && e && nrm && dz && M > 0 && F > 0, "tt_l2norm_backward: bad arguments");
  TT_CHECK_ARG(e_dtype == TT_F32 || e_dtype == TT_BF16, "tt_l2norm_backward: bad dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (e_dtype == TT_F32)
    l2norm_bwd_kernel<float><<<ceil_div(M, 8), 256, 0, s>>>(de, (const float*)e, nrm, M, F, dz);
  else
    l2norm_bwd_kernel<__nv_bfloat16><<<ceil_div(M, 8), 256, 0, s>>>(de, (const __nv_bfloat16*)e, nrm, M, F, dz);
  TT_LAUNCH_CHECK("l2norm_bwd_kernel");
  return TT_OK;
}

}  // extern "C"
