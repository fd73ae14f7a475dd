// Retrieval metrics without the B x B matrix: rank of every positive pair, recall@k, MRR.
//
// The reference's evaluator materialises S, runs topk twice and a full argsort, then walks the rows in a Python loop
// (src/evaluation/evaluator.py:20-71). All of it is a function of one integer per row:
//     rank_i = #{ j : S_ij > S_ii } + #{ j < i' : S_ij == S_ii }        (i' = the positive's column; stable descending order)
// recall@k = mean(rank_i < k), MRR = mean(1 / (rank_i + 1)). The logits are produced tile by tile in registers with the same
// fp32 main loop as tt_similarity_matrix (bit-identical values), compared against the row's positive logit and counted with
// integer atomics (order-independent, hence deterministic). Nothing of size B x B is stored.
#include "sgemm.cuh"

namespace tt {

// pass 1 (diagonal tiles only) writes diag[m] = S[m, diag_offset + m]; pass 2 (all tiles) counts
template <bool COUNT>
__global__ void __launch_bounds__(GTHREADS)
eval_rank_tile_kernel(const float* __restrict__ X, const float* __restrict__ Y, int64_t Bx, int64_t By, int F, float inv_temp,
                      int64_t diag_offset, float* __restrict__ diag, int32_t* __restrict__ rank) {
  __shared__ GemmSmem sm;
  int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
  if (!COUNT) {
    // blockIdx.x enumerates the (at most two) column tiles that hold this row block's positives
    n0 = ((m0 + diag_offset) / GN + blockIdx.x) * GN;
    if (n0 >= By) return;
  }
  float acc[4][4] = {};
  gemm_mainloop<true, true>(X, F, m0, Bx, Y, F, n0, By, 0, F, sm, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= Bx) continue;
    const int64_t pos = diag_offset + m;  // column of the positive
    if (!COUNT) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tx * 4 + j == pos) diag[m] = acc[i][j] * inv_temp;
    } else {
      const float d = diag[m];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (n < By && n != pos) {
          const float s = acc[i][j] * inv_temp;
          cnt += (s > d) || (s == d && n < pos);
        }
      }
      // the 16 lanes (tx) of a half-warp share row m
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (tx == 0 && cnt) atomicAdd(rank + m, cnt);
    }
  }
}

// rank from a materialised matrix (the reference evaluator's own entry points take S): one warp per row
__global__ void __launch_bounds__(256)
eval_rank_matrix_kernel(const float* __restrict__ S, int64_t Bx, int64_t By, int64_t lds, int64_t diag_offset,
                        int32_t* __restrict__ rank) {
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= Bx) return;
  const int lane = threadIdx.x & 31;
  const int64_t pos = diag_offset + m;
  const float* row = S + m * lds;
  const float d = row[pos];
  int cnt = 0;
  for (int64_t n = lane; n < By; n += 32) {
    const float s = __ldg(row + n);
    cnt += (n != pos) && ((s > d) || (s == d && n < pos));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) rank[m] = cnt;
}

// out[t] = mean(rank < ks[t]) for t < nk; out[nk] = mean(1/(rank+1)); out[nk+1] = mean(rank == 0) (top-1 accuracy).
// One CTA, fixed reduction order; integer hit counts, fp64 reciprocal-rank sum.
constexpr int EVAL_MAX_K = 8;
__global__ void __launch_bounds__(1024)
eval_rank_metrics_kernel(const int32_t* __restrict__ rank, int64_t B, const int32_t* __restrict__ ks, int nk,
                         float* __restrict__ out) {
  __shared__ long long s_hit[32][EVAL_MAX_K + 1];
  __shared__ double s_rr[32];
  long long hit[EVAL_MAX_K + 1] = {};
  double rr = 0.0;
  int kk[EVAL_MAX_K];
  for (int t = 0; t < EVAL_MAX_K; ++t) kk[t] = t < nk ? ks[t] : 0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const int r = rank[i];
#pragma unroll
    for (int t = 0; t < EVAL_MAX_K; ++t) hit[t] += (t < nk && r < kk[t]);
    hit[EVAL_MAX_K] += (r == 0);
    rr += 1.0 / (double)(r + 1);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int t = 0; t <= EVAL_MAX_K; ++t) hit[t] += __shfl_xor_sync(0xffffffffu, hit[t], o);
    rr += __shfl_xor_sync(0xffffffffu, rr, o);
  }
  if (lane == 0) {
    for (int t = 0; t <= EVAL_MAX_K; ++t) s_hit[warp][t] = hit[t];
    s_rr[warp] = rr;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    double tot = 0.0;
    for (int w = 0; w < nw; ++w) tot += s_rr[w];
    for (int t = 0; t <= EVAL_MAX_K; ++t) {
      long long h = 0;
      for (int w = 0; w < nw; ++w) h += s_hit[w][t];
      if (t < nk) out[t] = (float)((double)h / (double)B);
      if (t == EVAL_MAX_K) out[nk + 1] = (float)((double)h / (double)B);
    }
    out[nk] = (float)(tot / (double)B);
  }
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_eval_ranks(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                  int64_t diag_offset, float* diag, int32_t* rank, tt_stream_t stream) {
  TT_CHECK_ARG(X && Y && diag && rank && Bx > 0 && By > 0 && F > 0, "tt_eval_ranks: bad arguments");
  TT_CHECK_ARG(diag_offset >= 0 && diag_offset + Bx <= By, "tt_eval_ranks: positives [diag_offset, diag_offset+Bx) must lie inside the By columns");
  if (dtype != TT_F32) {
    set_error("tt_eval_ranks: dtype %d not supported (evaluation runs in fp32)", dtype);
    return TT_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TT_CUDA(cudaMemsetAsync(rank, 0, (size_t)Bx * sizeof(int32_t), s));
  const int nRT = ceil_div(Bx, GM);
  eval_rank_tile_kernel<false><<<dim3(2, nRT), GTHREADS, 0, s>>>((const float*)X, (const float*)Y, Bx, By, F, inv_temp, diag_offset,
                                                                 diag, rank);
  eval_rank_tile_kernel<true><<<dim3(ceil_div(By, GN), nRT), GTHREADS, 0, s>>>((const float*)X, (const float*)Y, Bx, By, F, inv_temp,
                                                                               diag_offset, diag, rank);
  TT_LAUNCH_CHECK("eval_rank_tile_kernel");
  return TT_OK;
}

int tt_eval_ranks_from_matrix(const float* S, int64_t Bx, int64_t By, int64_t lds, int64_t diag_offset, int32_t* rank,
                              tt_stream_t stream) {
  TT_CHECK_ARG(S && rank && Bx > 0 && By > 0 && lds >= By, "tt_eval_ranks_from_matrix: bad arguments");
  TT_CHECK_ARG(diag_offset >= 0 && diag_offset + Bx <= By, "tt_eval_ranks_from_matrix: positives outside the matrix");
  eval_rank_matrix_kernel<<<ceil_div(Bx, 8), 256, 0, (cudaStream_t)stream>>>(S, Bx, By, lds, diag_offset, rank);
  TT_LAUNCH_CHECK("eval_rank_matrix_kernel");
  return TT_OK;
}

int tt_eval_rank_metrics(const int32_t* rank, int64_t B, const int32_t* d_ks, int32_t nk, float* out, tt_stream_t stream) {
  TT_CHECK_ARG(rank && out && B > 0 && nk >= 0 && nk <= EVAL_MAX_K && (nk == 0 || d_ks), "tt_eval_rank_metrics: bad arguments (nk <= 8)");
  eval_rank_metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rank, B, d_ks, nk, out);
  TT_LAUNCH_CHECK("eval_rank_metrics_kernel");
  return TT_OK;
}

}  // extern "C"
