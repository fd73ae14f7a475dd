// In-batch-negative similarity logits + symmetric softmax cross-entropy, exact-fp32 SIMT path.
// The B x B logits are produced tile by tile in registers and reduced on the fly; they are never
// written to memory.  (The bf16 path does the same contraction on tcgen05 tensor cores with the
// tile held in TMEM: simce_tc.cu.)
#include "sgemm.cuh"
#include "simce_tc.h"
#include <float.h>

namespace tt {

// ------------------------------------------------------------------------------------------------
// forward: one 64x64 logits tile per CTA -> row/column partial reductions
// ------------------------------------------------------------------------------------------------
struct SimceFwdWs {
  float* row_e;    // [nCT][Bx] sum_j exp(s - shift)
  float* row_s;    // [nCT][Bx] sum_j s
  float* row_max;  // [nCT][Bx]
  int32_t* row_arg;
  float* col_e;  // [nRT][By]
  float* col_s;
};

__global__ void __launch_bounds__(GTHREADS)
simce_fwd_tile_kernel(const float* __restrict__ X, const float* __restrict__ Y, int64_t Bx, int64_t By, int F,
                      float inv_temp, int64_t diag_offset, SimceFwdWs ws, float* __restrict__ diag) {
  __shared__ GemmSmem sm;
  __shared__ float s_ce[16][GN], s_cs[16][GN];
  const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  float acc[4][4] = {};
  gemm_mainloop<true, true>(X, F, m0, Bx, Y, F, n0, By, 0, F, sm, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float ce[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    float re = 0.f, rs = 0.f, rmax = -FLT_MAX;
    int32_t rarg = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      const bool valid = (m < Bx) && (n < By);
      const float s = acc[i][j] * inv_temp;
      const float e = valid ? expf(s - inv_temp) : 0.f;
      if (valid) {
        re += e;
        rs += s;
        ce[j] += e;
        cs[j] += s;
        if (s > rmax) {  // ascending n: strict '>' keeps the first maximum
          rmax = s;
          rarg = (int32_t)n;
        }
        if (n == diag_offset + m) diag[m] = s;
      }
    }
    // reduce across the 16 lanes (tx) of this half-warp, fixed butterfly order
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      rs += __shfl_xor_sync(0xffffffffu, rs, o);
      const float om = __shfl_xor_sync(0xffffffffu, rmax, o);
      const int32_t oa = __shfl_xor_sync(0xffffffffu, rarg, o);
      if (om > rmax || (om == rmax && oa < rarg)) {
        rmax = om;
        rarg = oa;
      }
    }
    if (tx == 0 && m < Bx) {
      const int64_t o = (int64_t)blockIdx.x * Bx + m;
      ws.row_e[o] = re;
      ws.row_s[o] = rs;
      ws.row_max[o] = rmax;
      ws.row_arg[o] = rarg;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_ce[ty][tx * 4 + j] = ce[j];
    s_cs[ty][tx * 4 + j] = cs[j];
  }
  __syncthreads();
  if (threadIdx.x < GN) {
    const int64_t n = n0 + threadIdx.x;
    if (n < By) {
      float e = 0.f, s = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        e += s_ce[k][threadIdx.x];
        s += s_cs[k][threadIdx.x];
      }
      ws.col_e[(int64_t)blockIdx.y * By + n] = e;
      ws.col_s[(int64_t)blockIdx.y * By + n] = s;
    }
  }
}

__global__ void simce_rows_reduce_kernel(SimceFwdWs ws, int64_t Bx, int nCT, float* __restrict__ row_sumexp,
                                         float* __restrict__ row_sum_s, int32_t* __restrict__ row_argmax) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Bx) return;
  float e = 0.f, s = 0.f, mx = -FLT_MAX;
  int32_t arg = 0x7fffffff;
  for (int t = 0; t < nCT; ++t) {
    const int64_t o = (int64_t)t * Bx + m;
    e += ws.row_e[o];
    s += ws.row_s[o];
    const float tm = ws.row_max[o];
    const int32_t ta = ws.row_arg[o];
    if (tm > mx || (tm == mx && ta < arg)) {
      mx = tm;
      arg = ta;
    }
  }
  row_sumexp[m] = e;
  if (row_sum_s) row_sum_s[m] = s;
  if (row_argmax) row_argmax[m] = arg;
}

__global__ void simce_cols_reduce_kernel(SimceFwdWs ws, int64_t By, int nRT, float* __restrict__ col_sumexp,
                                         float* __restrict__ col_sum_s) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= By) return;
  float e = 0.f, s = 0.f;
  for (int t = 0; t < nRT; ++t) {
    e += ws.col_e[(int64_t)t * By + n];
    s += ws.col_s[(int64_t)t * By + n];
  }
  col_sumexp[n] = e;
  if (col_sum_s) col_sum_s[n] = s;
}

// ------------------------------------------------------------------------------------------------
// finalize: sums -> reciprocal sums for the backward + five partial scalars (deterministic tree)
// ------------------------------------------------------------------------------------------------
constexpr int FIN_BLOCKS = 16;  // partial is f64[FIN_BLOCKS][8]: block partials, summed in block order by simce_scalars_kernel

__global__ void __launch_bounds__(512)
simce_finalize_kernel(const float* __restrict__ row_sumexp, const float* __restrict__ col_sumexp,
                      const float* __restrict__ row_sum_s, const float* __restrict__ col_sum_s,
                      const float* __restrict__ diag, const int32_t* __restrict__ row_argmax, int64_t Bx, int64_t By,
                      int64_t diag_offset, double inv_bg, float inv_temp, float ls, float* __restrict__ inv_row,
                      float* __restrict__ inv_col, double* __restrict__ partial) {
  __shared__ double red[5][16];
  const int t = threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double a[5] = {0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + t; i < Bx; i += stride) {
    const float se = row_sumexp[i];
    inv_row[i] = 1.f / se;
    const double lse_r = (double)inv_temp + (double)logf(se);  // fp32 log (1 ulp) is plenty: the sum over rows is fp64
    const double d = (double)diag[i];
    const double rss = row_sum_s ? (double)row_sum_s[i] : 0.0;
    a[0] += lse_r - (1.0 - ls) * d - (double)ls * inv_bg * rss;
    const int64_t j = diag_offset + i;
    if (j < By) {
      const double lse_c = (double)inv_temp + (double)logf(col_sumexp[j]);
      const double css = col_sum_s ? (double)col_sum_s[j] : 0.0;
      a[1] += lse_c - (1.0 - ls) * d - (double)ls * inv_bg * css;
    }
    if (row_argmax && (int64_t)row_argmax[i] == j) a[2] += 1.0;
    a[3] += d;
    a[4] += rss;
  }
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + t; j < By; j += stride) inv_col[j] = 1.f / col_sumexp[j];
  // fixed-shape reduction: butterfly inside each warp, then warp 0 folds the 16 warp sums; blocks are folded in order later
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
  }
  if ((t & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) red[k][t >> 5] = a[k];
  }
  __syncthreads();
  if (t < 32) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double v = t < 16 ? red[k][t] : 0.0;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (t == 0) partial[blockIdx.x * 8 + k] = v;
    }
    if (t >= 5 && t < 8) partial[blockIdx.x * 8 + t] = 0.0;
  }
}

__global__ void simce_scalars_kernel(const double* __restrict__ part, long long Bg, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double p[5] = {0, 0, 0, 0, 0};
  for (int blk = 0; blk < FIN_BLOCKS; ++blk)
    for (int k = 0; k < 5; ++k) p[k] += part[blk * 8 + k];
  const double b = (double)Bg;
  const double loss = 0.5 * (p[0] + p[1]) / b;
  const double pos = p[3] / b;
  const double neg = Bg > 1 ? (p[4] - p[3]) / (b * b - b) : nan("");
  out[0] = (float)loss;
  out[1] = (float)(p[2] / b);
  out[2] = (float)pos;
  out[3] = (float)neg;
  out[4] = (float)(pos - neg);
}

// ------------------------------------------------------------------------------------------------
// backward: one CTA per 64-row block, loops over all column tiles, recomputes S, forms dS in
// registers and accumulates dX += dS · Y.  Called twice (rows = notice, rows = company).
// ------------------------------------------------------------------------------------------------
constexpr int SB = 64, SBP = SB + 4;

template <int CF>
__global__ void __launch_bounds__(256)
simce_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, int64_t Bx, int64_t By, int F, float inv_temp,
                 int64_t diag_offset, const float* __restrict__ inv_x, const float* __restrict__ inv_y, float gscale,
                 float tgt_diag, float tgt_off, float* __restrict__ dX) {
  extern __shared__ float smem[];
  float* Xt = smem;                 // [F][SBP]   Xt[f][i]
  float* Yt = Xt + (size_t)F * SBP;  // [F][SBP]   Yt[f][j]
  float* Yn = Yt + (size_t)F * SBP;  // [SB][F+4]  Yn[j][f]
  float* dSt = Yn + (size_t)SB * (F + 4);  // [SB][SBP]  dSt[j][i]
  const int FP = F + 4;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * SB;
  for (int idx = tid; idx < SB * F; idx += 256) {
    const int i = idx / F, f = idx - i * F;
    Xt[f * SBP + i] = (m0 + i < Bx) ? X[(m0 + i) * F + f] : 0.f;
  }
  float ix[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ix[i] = (m0 + ty * 4 + i < Bx) ? inv_x[m0 + ty * 4 + i] : 0.f;
  float acc2[4][CF];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CF; ++c) acc2[i][c] = 0.f;

  for (int64_t n0 = 0; n0 < By; n0 += SB) {
    __syncthreads();  // previous iteration's readers of Yt/Yn/dSt are done (also covers the Xt fill)
    for (int idx = tid; idx < SB * F; idx += 256) {
      const int j = idx / F, f = idx - j * F;
      const float v = (n0 + j < By) ? Y[(n0 + j) * F + f] : 0.f;
      Yt[f * SBP + j] = v;
      Yn[j * FP + f] = v;
    }
    __syncthreads();
    float acc[4][4] = {};
    for (int f = 0; f < F; ++f) {
      const float4 av = *reinterpret_cast<const float4*>(&Xt[f * SBP + ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Yt[f * SBP + tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      const float iy = n < By ? inv_y[n] : 0.f;
      float d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        const bool valid = (m < Bx) && (n < By);
        const float s = acc[i][j] * inv_temp;
        const float e = expf(s - inv_temp);
        const float tgt = (n == diag_offset + m) ? tgt_diag : tgt_off;
        d[i] = valid ? gscale * (e * (ix[i] + iy) - 2.f * tgt) : 0.f;
      }
      *reinterpret_cast<float4*>(&dSt[(tx * 4 + j) * SBP + ty * 4]) = make_float4(d[0], d[1], d[2], d[3]);
    }
    __syncthreads();
    for (int j = 0; j < SB; ++j) {
      const float4 av = *reinterpret_cast<const float4*>(&dSt[j * SBP + ty * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int c = 0; c < CF; ++c) {
        const int f = tx + 16 * c;
        const float b = f < F ? Yn[j * FP + f] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc2[i][c] = fmaf(a[i], b, acc2[i][c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= Bx) continue;
#pragma unroll
    for (int c = 0; c < CF; ++c) {
      const int f = tx + 16 * c;
      if (f < F) dX[m * F + f] = acc2[i][c];
    }
  }
}

static inline int64_t a256(int64_t x) { return (x + 255) / 256 * 256; }

static SimceFwdWs carve_ws(void* workspace, int64_t Bx, int64_t By) {
  const int64_t nCT = (By + GN - 1) / GN, nRT = (Bx + GM - 1) / GM;
  char* p = (char*)workspace;
  SimceFwdWs ws;
  ws.row_e = (float*)p;   p += a256(nCT * Bx * 4);
  ws.row_s = (float*)p;   p += a256(nCT * Bx * 4);
  ws.row_max = (float*)p; p += a256(nCT * Bx * 4);
  ws.row_arg = (int32_t*)p; p += a256(nCT * Bx * 4);
  ws.col_e = (float*)p;   p += a256(nRT * By * 4);
  ws.col_s = (float*)p;
  return ws;
}

}  // namespace tt

using namespace tt;

extern "C" {

int64_t tt_simce_workspace_bytes(int64_t Bx, int64_t By, int32_t F) {
  const int64_t nCT = (By + GN - 1) / GN, nRT = (Bx + GM - 1) / GM;
  const int64_t simt = 4 * a256(nCT * Bx * 4) + 2 * a256(nRT * By * 4) + 256;
  const int64_t tc = (F == 64 || F == 128) ? simce_tc_workspace_bytes(Bx, By, F) : 0;
  return simt > tc ? simt : tc;
}

int tt_debug_simce_timeline(void* d_buf_4x2048_i64) {
  simce_tc_set_timeline((long long*)d_buf_4x2048_i64);
  return TT_OK;
}

int tt_debug_simce_knobs(int32_t bits) {
  simce_tc_set_debug(bits);
  return TT_OK;
}

/* test hook: force the generic-proxy operand loader instead of TMA in the tcgen05 kernels */
int tt_debug_simce_force_manual_loads(int32_t on) {
  simce_tc_set_force_manual(on);
  return TT_OK;
}

int tt_simce_forward(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                     int64_t diag_offset, float* row_sumexp, float* col_sumexp, float* row_sum_s, float* col_sum_s,
                     float* diag, int32_t* row_argmax, void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(X && Y && row_sumexp && col_sumexp && diag && Bx > 0 && By > 0 && F > 0, "tt_simce_forward: bad arguments");
  TT_CHECK_ARG(diag_offset >= 0 && diag_offset + Bx <= By, "tt_simce_forward: diagonal [%lld,%lld) outside By=%lld",
               (long long)diag_offset, (long long)(diag_offset + Bx), (long long)By);
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_simce_workspace_bytes(Bx, By, F), "tt_simce_forward: workspace too small");
  TT_CHECK_ARG(inv_temp > 0.f && inv_temp <= 40.f, "tt_simce_forward: temperature must be >= 0.025 (inv_temp=%f)", inv_temp);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == TT_BF16) {
    if (F != 64 && F != 128) {
      set_error("tt_simce_forward: the tcgen05 bf16 path needs final_embedding_dim 64 or 128 (got %d)", F);
      return TT_ERR_UNSUPPORTED;
    }
    TT_CHECK_ARG(((uintptr_t)X & 15) == 0 && ((uintptr_t)Y & 15) == 0, "tt_simce_forward: bf16 operands must be 16-byte aligned");
    // rows of S (with metrics), then rows of S^T == columns of S
    int rc = simce_tc_rows(X, Y, Bx, By, F, inv_temp, diag_offset, row_sumexp, row_sum_s, diag, row_argmax, 1, workspace, s);
    if (rc != TT_OK) return rc;
    return simce_tc_rows(Y, X, By, Bx, F, inv_temp, 0, col_sumexp, col_sum_s, nullptr, nullptr, 0, workspace, s);
  }
  if (dtype != TT_F32) {
    set_error("tt_simce_forward: bad dtype %d", dtype);
    return TT_ERR_UNSUPPORTED;
  }
  SimceFwdWs ws = carve_ws(workspace, Bx, By);
  const int nCT = ceil_div(By, GN), nRT = ceil_div(Bx, GM);
  simce_fwd_tile_kernel<<<dim3(nCT, nRT), GTHREADS, 0, s>>>((const float*)X, (const float*)Y, Bx, By, F, inv_temp,
                                                            diag_offset, ws, diag);
  simce_rows_reduce_kernel<<<ceil_div(Bx, 256), 256, 0, s>>>(ws, Bx, nCT, row_sumexp, row_sum_s, row_argmax);
  simce_cols_reduce_kernel<<<ceil_div(By, 256), 256, 0, s>>>(ws, By, nRT, col_sumexp, col_sum_s);
  TT_LAUNCH_CHECK("simce_forward");
  return TT_OK;
}

/* Split forward for the train step (bf16 / tcgen05 path): what the BACKWARD needs — sum_j exp for rows and columns and their
 * reciprocals — is produced by tt_simce_forward_sums; what only the loss value and the metrics need (sum_j S_ij, arg-max) by
 * tt_simce_forward_metrics from the partials the first call left in the workspace, typically on a side stream next to the
 * backward pass. Together they write exactly what tt_simce_forward writes. */
int64_t tt_simce_split_workspace_bytes(int64_t Bx, int64_t By, int32_t F) {
  return (F == 64 || F == 128) ? 2 * a256(simce_tc_workspace_bytes(Bx, By, F)) : 0;
}

int tt_simce_forward_sums(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                          int64_t diag_offset, float* row_sumexp, float* col_sumexp, float* inv_row, float* inv_col, float* diag,
                          void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(X && Y && row_sumexp && col_sumexp && inv_row && inv_col && diag && Bx > 0 && By > 0, "tt_simce_forward_sums: bad arguments");
  TT_CHECK_ARG(dtype == TT_BF16 && (F == 64 || F == 128), "tt_simce_forward_sums: bf16 with final_embedding_dim 64 or 128 only");
  TT_CHECK_ARG(diag_offset >= 0 && diag_offset + Bx <= By, "tt_simce_forward_sums: diagonal outside By");
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_simce_split_workspace_bytes(Bx, By, F), "tt_simce_forward_sums: workspace too small");
  TT_CHECK_ARG(inv_temp > 0.f && inv_temp <= 40.f, "tt_simce_forward_sums: temperature must be >= 0.025");
  TT_CHECK_ARG(((uintptr_t)X & 15) == 0 && ((uintptr_t)Y & 15) == 0, "tt_simce_forward_sums: operands must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  char* w1 = (char*)workspace + a256(simce_tc_workspace_bytes(Bx, By, F));
  int rc = simce_tc_rows(X, Y, Bx, By, F, inv_temp, diag_offset, row_sumexp, nullptr, diag, nullptr, 1, workspace, s, 1, inv_row);
  if (rc != TT_OK) return rc;
  return simce_tc_rows(Y, X, By, Bx, F, inv_temp, 0, col_sumexp, nullptr, nullptr, nullptr, 0, w1, s, 1, inv_col);
}

int tt_simce_forward_metrics(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                             float* row_sum_s, float* col_sum_s, int32_t* row_argmax, void* workspace, int64_t workspace_bytes,
                             tt_stream_t stream) {
  TT_CHECK_ARG(X && Y && row_sum_s && row_argmax && Bx > 0 && By > 0, "tt_simce_forward_metrics: bad arguments");
  TT_CHECK_ARG(dtype == TT_BF16 && (F == 64 || F == 128), "tt_simce_forward_metrics: bf16 with final_embedding_dim 64 or 128 only");
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_simce_split_workspace_bytes(Bx, By, F), "tt_simce_forward_metrics: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  char* w1 = (char*)workspace + a256(simce_tc_workspace_bytes(Bx, By, F));
  int rc = simce_tc_rows_metrics(X, Y, Bx, By, F, inv_temp, row_sum_s, row_argmax, workspace, s);
  if (rc != TT_OK || !col_sum_s) return rc;
  return simce_tc_rows_metrics(Y, X, By, Bx, F, inv_temp, col_sum_s, nullptr, w1, s);
}

/* ---- three-pass train step of the bf16 path (pre-scaled operands; simce_tc.cu) ---- */
int64_t tt_simce3_workspace_bytes(int64_t Bx, int64_t By, int32_t F) {
  if (F != 64 && F != 128) return 0;
  return a256(simce_tc_workspace_bytes(Bx, By, F)) + simce_tc_grad_workspace_bytes(Bx, By, F);
}
static char* simce3_grad_ws(void* workspace, int64_t Bx, int64_t By, int F) {
  return (char*)workspace + a256(simce_tc_workspace_bytes(Bx, By, F));
}
#define SIMCE3_COMMON(name)                                                                                                  \
  TT_CHECK_ARG(Bx > 0 && By > 0 && (F == 64 || F == 128), name ": needs final_embedding_dim 64 or 128");                     \
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_simce3_workspace_bytes(Bx, By, F), name ": workspace too small");          \
  TT_CHECK_ARG(inv_temp > 0.f && inv_temp <= 40.f, name ": temperature must be >= 0.025");

int tt_simce3_rows(const void* Xs, const void* Y, int64_t Bx, int64_t By, int32_t F, float inv_temp, int64_t diag_offset, float* diag,
                   void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  SIMCE3_COMMON("tt_simce3_rows")
  TT_CHECK_ARG(Xs && Y && diag && diag_offset >= 0 && diag_offset + Bx <= By, "tt_simce3_rows: bad arguments");
  TT_CHECK_ARG(((uintptr_t)Xs & 15) == 0 && ((uintptr_t)Y & 15) == 0, "tt_simce3_rows: operands must be 16-byte aligned");
  return simce_tc_rows_pre(Xs, Y, Bx, By, F, inv_temp, diag_offset, diag, workspace, (cudaStream_t)stream);
}

int tt_simce3_rows_finish(void* workspace, int64_t workspace_bytes, int64_t Bx, int64_t By, int32_t F, float inv_temp,
                          float* row_sumexp, float* inv_row, const void* X_unit, float gscale, void* Xs_out, tt_stream_t stream) {
  SIMCE3_COMMON("tt_simce3_rows_finish")
  TT_CHECK_ARG(row_sumexp && inv_row && (!Xs_out || X_unit), "tt_simce3_rows_finish: bad arguments");
  return simce_tc_rows_finish(workspace, Bx, By, F, inv_temp, row_sumexp, inv_row, X_unit, gscale, Xs_out, (cudaStream_t)stream);
}

int tt_simce3_grad(const void* Xs, const void* Y, const void* Ys, int64_t Bx, int64_t By, int32_t F, float inv_temp,
                   int64_t diag_offset, const float* inv_x, float gscale, float label_smoothing, void* workspace,
                   int64_t workspace_bytes, tt_stream_t stream) {
  SIMCE3_COMMON("tt_simce3_grad")
  TT_CHECK_ARG(Xs && Y && Ys && diag_offset >= 0 && diag_offset + Bx <= By, "tt_simce3_grad: bad arguments");
  TT_CHECK_ARG(((uintptr_t)Xs & 15) == 0 && ((uintptr_t)Y & 15) == 0 && ((uintptr_t)Ys & 15) == 0, "tt_simce3_grad: operands must be 16-byte aligned");
  return simce_tc_grad_pre(Xs, Y, Ys, Bx, By, F, inv_temp, diag_offset, inv_x, gscale, label_smoothing != 0.f,
                           simce3_grad_ws(workspace, Bx, By, F), (cudaStream_t)stream);
}

int tt_simce3_grad_finish(void* workspace, int64_t workspace_bytes, int64_t Bx, int64_t By, int32_t F, float inv_temp,
                          int64_t diag_offset, const void* Y, int32_t deferred, float gscale, float label_smoothing, int64_t B_global,
                          float* dX, float* sumexp_out, float* inv_out, const void* X_unit, void* Xs_out, tt_stream_t stream) {
  SIMCE3_COMMON("tt_simce3_grad_finish")
  TT_CHECK_ARG(Y && dX && B_global > 0 && (!Xs_out || (X_unit && deferred)), "tt_simce3_grad_finish: bad arguments");
  const float t_off = label_smoothing / (float)B_global;
  return simce_tc_grad_finish(simce3_grad_ws(workspace, Bx, By, F), Bx, By, F, inv_temp, diag_offset, Y, deferred, gscale,
                              (1.f - label_smoothing) + t_off, t_off, dX, sumexp_out, inv_out, X_unit, Xs_out, (cudaStream_t)stream);
}

int tt_simce3_scale_rows(const void* Y_unit, const float* sumexp, float gscale, float inv_temp, int64_t By, int32_t F, void* Ys_out,
                         tt_stream_t stream) {
  TT_CHECK_ARG(Y_unit && sumexp && Ys_out && By > 0 && (F == 64 || F == 128), "tt_simce3_scale_rows: bad arguments");
  return simce_tc_scale_rows(Y_unit, sumexp, gscale, inv_temp, By, F, Ys_out, (cudaStream_t)stream);
}

int tt_simce3_metrics(const void* Xs, const void* Y, int64_t Bx, int64_t By, int32_t F, float* row_sum_s, int32_t* row_argmax,
                      void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(Xs && Y && Bx > 0 && By > 0 && (F == 64 || F == 128) && (row_sum_s || row_argmax), "tt_simce3_metrics: bad arguments");
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_simce3_workspace_bytes(Bx, By, F), "tt_simce3_metrics: workspace too small");
  /* Xs carries c1 = log2(e)/T: X'.y * ln 2 = S / T */
  return simce_tc_rows_metrics(Xs, Y, Bx, By, F, 0.6931471805599453f, row_sum_s, row_argmax, workspace, (cudaStream_t)stream);
}

int tt_simce_finalize(const float* row_sumexp, const float* col_sumexp, const float* row_sum_s,
                      const float* col_sum_s, const float* diag, const int32_t* row_argmax, int64_t Bx, int64_t By,
                      int64_t diag_offset, int64_t B_global, float inv_temp, float label_smoothing, float* inv_row,
                      float* inv_col, double* partial, tt_stream_t stream) {
  TT_CHECK_ARG(row_sumexp && col_sumexp && diag && inv_row && inv_col && partial && Bx > 0 && By > 0 && B_global > 0,
               "tt_simce_finalize: bad arguments");
  TT_CHECK_ARG(label_smoothing == 0.f || (row_sum_s && col_sum_s), "tt_simce_finalize: label smoothing needs the sum_s vectors");
  simce_finalize_kernel<<<FIN_BLOCKS, 512, 0, (cudaStream_t)stream>>>(row_sumexp, col_sumexp, row_sum_s, col_sum_s, diag,
                                                             row_argmax, Bx, By, diag_offset, 1.0 / (double)B_global,
                                                             inv_temp, label_smoothing, inv_row, inv_col, partial);
  TT_LAUNCH_CHECK("simce_finalize_kernel");
  return TT_OK;
}

int tt_simce_scalars(const double* partial, int64_t B_global, float* out5, tt_stream_t stream) {
  TT_CHECK_ARG(partial && out5 && B_global > 0, "tt_simce_scalars: bad arguments");
  simce_scalars_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partial, (long long)B_global, out5);
  TT_LAUNCH_CHECK("simce_scalars_kernel");
  return TT_OK;
}

int tt_simce_backward(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                      int64_t diag_offset, const float* inv_x, const float* inv_y, float gscale,
                      float label_smoothing, int64_t B_global, float* dX, void* workspace, int64_t workspace_bytes,
                      tt_stream_t stream) {
  TT_CHECK_ARG(X && Y && inv_x && inv_y && dX && Bx > 0 && By > 0 && F > 0 && B_global > 0, "tt_simce_backward: bad arguments");
  if (dtype == TT_BF16) {
    if (F != 64 && F != 128) {
      set_error("tt_simce_backward: the tcgen05 bf16 path needs final_embedding_dim 64 or 128 (got %d)", F);
      return TT_ERR_UNSUPPORTED;
    }
    TT_CHECK_ARG(workspace && workspace_bytes >= simce_tc_workspace_bytes(Bx, By, F), "tt_simce_backward: workspace too small");
    const float t_off = label_smoothing / (float)B_global;
    return simce_tc_bwd(X, Y, Bx, By, F, inv_temp, diag_offset, inv_x, inv_y, gscale, (1.f - label_smoothing) + t_off, t_off, dX,
                        workspace, (cudaStream_t)stream);
  }
  if (dtype != TT_F32) {
    set_error("tt_simce_backward: bad dtype %d", dtype);
    return TT_ERR_UNSUPPORTED;
  }
  if (F > 128) {
    set_error("tt_simce_backward: final_embedding_dim %d > 128 not supported", F);
    return TT_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const float tgt_off = label_smoothing / (float)B_global;
  const float tgt_diag = (1.f - label_smoothing) + tgt_off;
  const size_t smem = ((size_t)2 * F * SBP + (size_t)SB * (F + 4) + (size_t)SB * SBP) * sizeof(float);
  const int grid = ceil_div(Bx, SB);
  if (F <= 64) {
    TT_CUDA(cudaFuncSetAttribute(simce_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    simce_bwd_kernel<4><<<grid, 256, smem, s>>>((const float*)X, (const float*)Y, Bx, By, F, inv_temp, diag_offset, inv_x,
                                                inv_y, gscale, tgt_diag, tgt_off, dX);
  } else {
    TT_CUDA(cudaFuncSetAttribute(simce_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    simce_bwd_kernel<8><<<grid, 256, smem, s>>>((const float*)X, (const float*)Y, Bx, By, F, inv_temp, diag_offset, inv_x,
                                                inv_y, gscale, tgt_diag, tgt_off, dX);
  }
  TT_LAUNCH_CHECK("simce_bwd_kernel");
  return TT_OK;
}

int tt_similarity_matrix(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype,
                         float inv_temp, float* S, int64_t lds, tt_stream_t stream) {
  TT_CHECK_ARG(X && Y && S && Bx > 0 && By > 0 && F > 0 && lds >= By, "tt_similarity_matrix: bad arguments");
  if (dtype != TT_F32) {
    set_error("tt_similarity_matrix: dtype %d not supported by the fp32 path", dtype);
    return TT_ERR_UNSUPPORTED;
  }
  dim3 grid(ceil_div(By, GN), ceil_div(Bx, GM), 1);
  sgemm_kernel<true, true><<<grid, GTHREADS, 0, (cudaStream_t)stream>>>((const float*)X, F, (const float*)Y, F, nullptr,
                                                                        S, lds, Bx, By, F, F, 0, inv_temp, 0);
  TT_LAUNCH_CHECK("similarity_matrix");
  return TT_OK;
}

}  // extern "C"
