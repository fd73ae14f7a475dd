// Tower Linear layers on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
// One generic kernel computes a 128 x TN output tile of  C[m,n] = sum_r A(m,r) * B(n,r)  for any combination of
// operand layouts, so forward, dgrad and wgrad of nn.Linear (base_tower.py:59,89,97 and their autograd) all map to it:
//   forward  y = x W^T + b : A = x  [M,K]  K-major      B = W  [N,K] K-major     reduce over K
//   dgrad    dx = dy W     : A = dy [M,N]  K-major      B = W  [N,K] MN-major    reduce over N
//   wgrad    dW^T = x^T dy : A = x  [M,K]  MN-major     B = dy [M,N] MN-major    reduce over the batch M (split-K)
// Operands stream through a 4-stage TMA ring (128B-swizzled boxes, 64 reduction elements per stage); one elected
// thread issues tcgen05.mma; four epilogue warps read the accumulator out of TMEM (one output row per thread) and
// apply bias / ReLU / dtype conversion, or write a transposed fp32 partial for the split-K weight gradient.
#include "tc_common.cuh"

namespace tt {

constexpr int GT_STAGES = 4;
constexpr int GT_THREADS = 192;  // warps 0-3 epilogue, 4 MMA issuer + TMEM allocator, 5 TMA producer
constexpr int GT_BK = 64;        // reduction elements per stage (one 128-byte swizzle span)

enum { EPI_STORE = 0, EPI_PARTIAL_T = 1 };

struct GemmTcParams {
  int64_t M, N, R;        // output rows, output cols, reduction length
  int tn;                 // output tile width: 64 or 128
  int a_mn, b_mn;         // operand major-ness (0 = reduction-contiguous "K-major", 1 = MN-major)
  int stages_per_split;   // reduction stages handled by one CTA (grid.z splits)
  int ring;               // smem ring depth = min(GT_STAGES, stages per CTA): short reductions leave room for 2-3 CTAs per SM
  int epi;
  // EPI_STORE
  void* C;
  int64_t ldc;
  int out_bf16;
  const float* bias;
  int relu;
  float* colstat;         // optional [gridDim.x][2][N]: per row tile, column sums of the STORED values and of their squares
                          // (BatchNorm statistics of the layer output, taken where the tile is still on chip)
  // EPI_PARTIAL_T: part[(split * N + n) * M + m]
  float* part;
};

struct GemmTcSmem {
  uint64_t full[GT_STAGES], empty[GT_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(GT_THREADS, 3)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmTcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = ATOM_BYTES;                 // 128 x 64 bf16 either way (1 K-major atom or 2 MN-major half-atoms)
  const int b_bytes = p.tn * 128;                 // tn x 64 bf16
  const int stage_bytes = a_bytes + b_bytes;
  const int ring = p.ring;
  GemmTcSmem* sb = reinterpret_cast<GemmTcSmem*>(base + ring * stage_bytes);
  // warp index through a shuffle: role dispatch and everything derived from kernel parameters stays warp-uniform, so the
  // tcgen05 / TMA operands live in uniform registers (a divergent `if (lane == 0)` makes ptxas wrap each one in a waterfall loop)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 128, n0 = (int64_t)blockIdx.y * p.tn;
  const int total_stages = (int)((p.R + GT_BK - 1) / GT_BK);
  const int s_begin = blockIdx.z * p.stages_per_split;
  const int s_end = min(s_begin + p.stages_per_split, total_stages);
  const int n_stages = s_end - s_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) {
      mbar_init(&sb->full[i], 1);
      mbar_init(&sb->empty[i], 1);
    }
    mbar_init(&sb->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sb->tmem_base, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  pdl_wait();  // programmatic dependent launch: the set-up above overlapped the predecessor's tail
  pdl_launch_dependents();

  if (warp == 5) {
    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_stages; ++i) {
      mbar_wait(&sb->empty[st], ph ^ 1);
      unsigned char* sa = base + st * stage_bytes;
      unsigned char* sbm = sa + a_bytes;
      const int r0 = (s_begin + i) * GT_BK;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&sb->full[st], (uint32_t)stage_bytes);
        if (!p.a_mn) {
          tma_load_2d(sa, &map_a, r0, (int)m0, &sb->full[st]);              // box {64 r, 128 m}
        } else {
          tma_load_2d(sa, &map_a, (int)m0, r0, &sb->full[st]);              // box {64 m, 64 r} x 2
          tma_load_2d(sa + 8192, &map_a, (int)m0 + 64, r0, &sb->full[st]);
        }
        if (!p.b_mn) {
          tma_load_2d(sbm, &map_b, r0, (int)n0, &sb->full[st]);             // box {64 r, tn n}
        } else {
          tma_load_2d(sbm, &map_b, (int)n0, r0, &sb->full[st]);             // box {64 n, 64 r} x (tn / 64)
          if (p.tn == 128) tma_load_2d(sbm + 8192, &map_b, (int)n0 + 64, r0, &sb->full[st]);
        }
      }
      __syncwarp();
      if (++st == ring) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 4) {
    const uint32_t idesc = make_idesc(128, p.tn, p.a_mn, p.b_mn);
    const uint32_t base_a = smem_u32(base);
    // descriptors of stage 0, k-step 0; stages and k-steps only move the 14-bit start-address field
    const uint64_t da0 = p.a_mn ? make_desc(base_a, 8192, 1024) : make_desc(base_a, 16, 1024);
    const uint64_t db0 = p.b_mn ? make_desc(base_a + a_bytes, 8192, 1024) : make_desc(base_a + a_bytes, 16, 1024);
    const uint32_t ka = p.a_mn ? (2048 >> 4) : (32 >> 4), kb = p.b_mn ? (2048 >> 4) : (32 >> 4);
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_stages; ++i) {
      mbar_wait(&sb->full[st], ph);
      tc_fence_after();
      const uint64_t da = da0 + (uint64_t)(st * stage16), db = db0 + (uint64_t)(st * stage16);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k) umma_bf16(tmem, da + (uint64_t)(k * ka), db + (uint64_t)(k * kb), idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&sb->empty[st]);
        if (i == n_stages - 1) umma_commit(&sb->acc_full);
      }
      __syncwarp();
      if (++st == ring) {
        st = 0;
        ph ^= 1;
      }
    }
  } else {
    // epilogue: thread == output row
    const int r_in = warp * 32 + lane;
    const int64_t m = m0 + r_in;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    if (n_stages > 0) {
      mbar_wait(&sb->acc_full, 0);
      tc_fence_after();
    }
    for (int c32 = 0; c32 < p.tn / 32; ++c32) {
      uint32_t r[32];
      if (n_stages > 0) {
        tmem_ld32(tmem + lane_addr + c32 * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int64_t nb = n0 + c32 * 32;
      if (p.epi == EPI_STORE) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]);
            if (p.bias && nb + j < p.N) x += __ldg(p.bias + nb + j);
            if (p.relu) x = fmaxf(x, 0.f);
            v[j] = x;
          }
          if (p.colstat) {  // what the consumer will read: the bf16-rounded value
#pragma unroll
            for (int j = 0; j < 32; ++j) if (p.out_bf16) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
          }
          if (nb + 32 <= p.N) {
            if (p.out_bf16) {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.C) + m * p.ldc + nb;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 w;
                const uint2 a = pack_bf16x4(make_float4(v[q * 8], v[q * 8 + 1], v[q * 8 + 2], v[q * 8 + 3]));
                const uint2 b = pack_bf16x4(make_float4(v[q * 8 + 4], v[q * 8 + 5], v[q * 8 + 6], v[q * 8 + 7]));
                w.x = a.x; w.y = a.y; w.z = b.x; w.w = b.y;
                *reinterpret_cast<uint4*>(o + q * 8) = w;
              }
            } else {
              float* o = reinterpret_cast<float*>(p.C) + m * p.ldc + nb;
#pragma unroll
              for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(o + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
          } else {
            for (int j = 0; j < 32 && nb + j < p.N; ++j) {
              if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.C)[m * p.ldc + nb + j] = __float2bfloat16_rn(v[j]);
              else reinterpret_cast<float*>(p.C)[m * p.ldc + nb + j] = v[j];
            }
          }
        }
        if (p.colstat) {
          // column sums over the tile's 128 rows through shared memory (the operand ring is free: every MMA has retired):
          // red[row][33] <- v; 4 x 32 partial sums; fixed order -> deterministic. Rows past M hold zeros.
          float* red = reinterpret_cast<float*>(base);
          float* red2 = red + 128 * 33;
#pragma unroll
          for (int j = 0; j < 32; ++j) red[r_in * 33 + j] = v[j];
          asm volatile("bar.sync 1, 128;" ::: "memory");
          {
            const int col = r_in & 31, q = r_in >> 5;
            float s1 = 0.f, s2 = 0.f;
            for (int r = q * 32; r < q * 32 + 32; ++r) {
              const float x = red[r * 33 + col];
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
            red2[q * 32 + col] = s1;
            red2[128 + q * 32 + col] = s2;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (r_in < 64) {
            const int which = r_in >> 5, col = r_in & 31;
            const float* src = red2 + which * 128 + col;
            const float t = (src[0] + src[32]) + (src[64] + src[96]);
            if (nb + col < p.N) p.colstat[((int64_t)blockIdx.x * 2 + which) * p.N + nb + col] = t;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      } else {  // transposed fp32 partial: consecutive threads (m) -> consecutive addresses
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) p.part[((int64_t)blockIdx.z * p.N + nb + j) * p.M + m] = __uint_as_float(r[j]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

__global__ void gt_splitk_reduce_kernel(const float* __restrict__ part, int64_t stride, int splits, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  int z = 0;
  for (; z + 8 <= splits; z += 8) {  // eight split partials in flight, added in split order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(z + u) * stride + i];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; z < splits; ++z) s += part[(int64_t)z * stride + i];
  out[i] = s;
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      *reinterpret_cast<uint2*>(dst + i) = pack_bf16x4(*reinterpret_cast<const float4*>(src + i));
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
  }
}

// several fp32 -> bf16 casts in ONE launch (blockIdx.y = segment): at the start of the captured train step every extra root
// node costs 5-9 us of launch ramp, so the weight mirror and the two dense inputs are cast together
struct CastSegs {
  const float* src[4];
  __nv_bfloat16* dst[4];
  int64_t n[4];
};
__global__ void __launch_bounds__(256) cast_bf16_multi_kernel(const CastSegs segs) {
  const float* __restrict__ src = segs.src[blockIdx.y];
  __nv_bfloat16* __restrict__ dst = segs.dst[blockIdx.y];
  const int64_t n = segs.n[blockIdx.y];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      *reinterpret_cast<uint2*>(dst + i) = pack_bf16x4(*reinterpret_cast<const float4*>(src + i));
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
  }
}

// column sums of a bf16 matrix (bias gradients), fp64 partials like the fp32 path
constexpr int GCS_ROWS = 64;
__global__ void __launch_bounds__(256)
colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ X, int64_t ld, int64_t M, int N, double* __restrict__ part) {
  __shared__ double s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int64_t m_lo = (int64_t)blockIdx.y * GCS_ROWS;
  const int64_t m_hi = (m_lo + GCS_ROWS < M) ? m_lo + GCS_ROWS : M;
  double a = 0.0;
  if (n < N)
    for (int64_t m = m_lo + ty; m < m_hi; m += 8) a += (double)__bfloat162float(X[m * ld + n]);
  s[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && n < N) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][tx];
    part[(int64_t)blockIdx.y * N + n] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_bf16_final_kernel(const double* __restrict__ part, int chunks, int N, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per column, fixed shuffle tree
  if (n >= N) return;
  double t = 0.0;
  for (int c = lane; c < chunks; c += 32) t += part[(int64_t)c * N + n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) out[n] = (float)t;
}

static size_t gt_smem_bytes(int tn, int ring) { return 1024 + (size_t)ring * (ATOM_BYTES + tn * 128) + sizeof(GemmTcSmem) + 64; }

static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, GemmTcParams& p, int splits, cudaStream_t s) {
  const int total_stages = ceil_div(p.R, (int64_t)GT_BK);
  const int per_cta = p.stages_per_split < total_stages ? p.stages_per_split : total_stages;
  p.ring = per_cta < GT_STAGES ? (per_cta < 1 ? 1 : per_cta) : GT_STAGES;
  const size_t smem = gt_smem_bytes(p.tn, p.ring);
  TT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem_bytes(128, GT_STAGES)));
  dim3 grid(ceil_div(p.M, 128), ceil_div(p.N, p.tn), splits);
  TT_CUDA(launch_pdl(gemm_tc_kernel, grid, dim3(GT_THREADS), smem, s, ma, mb, p));
  TT_LAUNCH_CHECK("gemm_tc_kernel");
  return TT_OK;
}

static bool aligned_ok(const void* p, int64_t ld) { return (((uintptr_t)p) & 15) == 0 && (ld % 8) == 0; }

static int wgrad_splits_tc(int64_t M, int32_t N, int32_t K) {
  const int tiles = ceil_div(K, 128) * ceil_div(N, N > 64 ? 128 : 64);
  const int total_stages = ceil_div(M, GT_BK);
  int want = (sm_count() + tiles - 1) / tiles;
  if (want > total_stages / 2) want = total_stages / 2;  // at least two stages per split
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return want;
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_cast_bf16(const float* src, void* dst, int64_t n, tt_stream_t stream) {
  TT_CHECK_ARG(src && dst && n >= 0, "tt_cast_bf16: bad arguments");
  TT_CHECK_ARG((((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 7) == 0, "tt_cast_bf16: buffers must be 16/8-byte aligned");
  if (n == 0) return TT_OK;
  int64_t want = ceil_div(n, 1024);
  const int64_t cap = (int64_t)sm_count() * 8;
  cast_bf16_kernel<<<(int)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  TT_LAUNCH_CHECK("cast_bf16_kernel");
  return TT_OK;
}

int tt_cast_bf16_multi(const float* src0, void* dst0, int64_t n0, const float* src1, void* dst1, int64_t n1, const float* src2,
                       void* dst2, int64_t n2, const float* src3, void* dst3, int64_t n3, tt_stream_t stream) {
  CastSegs segs = {};
  const float* srcs[4] = {src0, src1, src2, src3};
  void* dsts[4] = {dst0, dst1, dst2, dst3};
  const int64_t ns[4] = {n0, n1, n2, n3};
  int cnt = 0;
  int64_t n_max = 0;
  for (int i = 0; i < 4; ++i) {
    if (ns[i] <= 0) continue;
    TT_CHECK_ARG(srcs[i] && dsts[i], "tt_cast_bf16_multi: NULL segment %d", i);
    TT_CHECK_ARG((((uintptr_t)srcs[i]) & 15) == 0 && (((uintptr_t)dsts[i]) & 7) == 0, "tt_cast_bf16_multi: segment %d is misaligned", i);
    segs.src[cnt] = srcs[i];
    segs.dst[cnt] = (__nv_bfloat16*)dsts[i];
    segs.n[cnt] = ns[i];
    n_max = ns[i] > n_max ? ns[i] : n_max;
    ++cnt;
  }
  if (cnt == 0) return TT_OK;
  int64_t want = ceil_div(n_max, 1024);
  const int64_t cap = (int64_t)sm_count() * 4;
  cast_bf16_multi_kernel<<<dim3((unsigned)(want > cap ? cap : want), cnt), 256, 0, (cudaStream_t)stream>>>(segs);
  TT_LAUNCH_CHECK("cast_bf16_multi_kernel");
  return TT_OK;
}

static int linear_forward_bf16_impl(const void* A, int64_t lda, const void* W_bf16, const float* bias, void* C, int64_t ldc,
                                    int32_t out_dtype, int64_t M, int32_t N, int32_t K, int32_t relu, float* colstat,
                                    tt_stream_t stream) {
  TT_CHECK_ARG(A && W_bf16 && C && M > 0 && N > 0 && K > 0 && lda >= K && ldc >= N, "tt_linear_forward_bf16: bad arguments");
  TT_CHECK_ARG(aligned_ok(A, lda) && aligned_ok(W_bf16, K), "tt_linear_forward_bf16: operands need 16-byte alignment and ld %% 8 == 0");
  TT_CHECK_ARG(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_linear_forward_bf16: bad out dtype");
  TT_CHECK_ARG((ldc % (out_dtype == TT_BF16 ? 8 : 4)) == 0 && (((uintptr_t)C) & 15) == 0,
               "tt_linear_forward_bf16: C needs 16-byte alignment and ldc %% 8 (bf16) / %% 4 (f32) == 0");
  GemmTcParams p = {};
  p.M = M; p.N = N; p.R = K; p.tn = N > 64 ? 128 : 64; p.a_mn = 0; p.b_mn = 0;
  p.stages_per_split = ceil_div(K, GT_BK); p.epi = EPI_STORE; p.C = C; p.ldc = ldc; p.out_bf16 = out_dtype == TT_BF16;
  p.bias = bias; p.relu = relu; p.colstat = colstat;
  CUtensorMap ma, mb;
  if (!make_map_bf16(&ma, A, M, K, lda, 64, 128) || !make_map_bf16(&mb, W_bf16, N, K, K, 64, p.tn)) {
    set_error("tt_linear_forward_bf16: cuTensorMapEncodeTiled failed");
    return TT_ERR_CUDA;
  }
  return launch_gemm(ma, mb, p, 1, (cudaStream_t)stream);
}

int tt_linear_forward_bf16(const void* A, int64_t lda, const void* W_bf16, const float* bias, void* C, int64_t ldc,
                           int32_t out_dtype, int64_t M, int32_t N, int32_t K, int32_t relu, tt_stream_t stream) {
  return linear_forward_bf16_impl(A, lda, W_bf16, bias, C, ldc, out_dtype, M, N, K, relu, nullptr, stream);
}

int64_t tt_linear_stats_tiles(int64_t M) { return (M + 127) / 128; }

int tt_linear_forward_bf16_stats(const void* A, int64_t lda, const void* W_bf16, const float* bias, void* C, int64_t ldc,
                                 int32_t out_dtype, int64_t M, int32_t N, int32_t K, int32_t relu, float* colstat,
                                 tt_stream_t stream) {
  TT_CHECK_ARG(colstat, "tt_linear_forward_bf16_stats: colstat is NULL");
  return linear_forward_bf16_impl(A, lda, W_bf16, bias, C, ldc, out_dtype, M, N, K, relu, colstat, stream);
}

int tt_linear_dgrad_bf16(const void* dY, int64_t ldy, const void* W_bf16, void* dX, int64_t ldx, int32_t out_dtype, int64_t M,
                         int32_t N, int32_t K, tt_stream_t stream) {
  TT_CHECK_ARG(dY && W_bf16 && dX && M > 0 && N > 0 && K > 0 && ldy >= N && ldx >= K, "tt_linear_dgrad_bf16: bad arguments");
  TT_CHECK_ARG(aligned_ok(dY, ldy) && aligned_ok(W_bf16, K), "tt_linear_dgrad_bf16: operands need 16-byte alignment and ld %% 8 == 0");
  TT_CHECK_ARG(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_linear_dgrad_bf16: bad out dtype");
  TT_CHECK_ARG((ldx % 8) == 0 && (((uintptr_t)dX) & 15) == 0, "tt_linear_dgrad_bf16: dX needs 16-byte alignment and ldx %% 8 == 0");
  GemmTcParams p = {};
  p.M = M; p.N = K; p.R = N; p.tn = K > 64 ? 128 : 64; p.a_mn = 0; p.b_mn = 1;
  p.stages_per_split = ceil_div(N, GT_BK); p.epi = EPI_STORE; p.C = dX; p.ldc = ldx; p.out_bf16 = out_dtype == TT_BF16;
  CUtensorMap ma, mb;
  // B(n_l = k, r = n) = W[n*K + k]: MN-major, the matrix is [N rows (reduction), K cols (output index)]
  if (!make_map_bf16(&ma, dY, M, N, ldy, 64, 128) || !make_map_bf16(&mb, W_bf16, N, K, K, 64, 64)) {
    set_error("tt_linear_dgrad_bf16: cuTensorMapEncodeTiled failed");
    return TT_ERR_CUDA;
  }
  return launch_gemm(ma, mb, p, 1, (cudaStream_t)stream);
}

int64_t tt_linear_wgrad_bf16_workspace_bytes(int64_t M, int32_t N, int32_t K) {
  const int64_t chunks = (M + GCS_ROWS - 1) / GCS_ROWS + 1;
  return (int64_t)wgrad_splits_tc(M, N, K) * N * K * 4 + chunks * N * 8 + 512;
}

int tt_linear_wgrad_bf16(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, float* dbias, int64_t M, int32_t N,
                         int32_t K, void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(dY && X && dW && M > 0 && N > 0 && K > 0 && ldy >= N && ldx >= K, "tt_linear_wgrad_bf16: bad arguments");
  TT_CHECK_ARG(aligned_ok(dY, ldy) && aligned_ok(X, ldx), "tt_linear_wgrad_bf16: operands need 16-byte alignment and ld %% 8 == 0");
  TT_CHECK_ARG(workspace && workspace_bytes >= tt_linear_wgrad_bf16_workspace_bytes(M, N, K), "tt_linear_wgrad_bf16: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int splits = wgrad_splits_tc(M, N, K);
  const int total_stages = ceil_div(M, GT_BK);
  const int sps = (total_stages + splits - 1) / splits;
  const int eff = (total_stages + sps - 1) / sps;
  // dW^T[k, n] = sum_m X[m,k] * dY[m,n]: logical M = K (in features), N = N (out features), both operands MN-major
  GemmTcParams p = {};
  p.M = K; p.N = N; p.R = M; p.tn = N > 64 ? 128 : 64; p.a_mn = 1; p.b_mn = 1;
  p.stages_per_split = sps; p.epi = EPI_PARTIAL_T;
  float* part = (float*)workspace;
  p.part = eff == 1 ? dW : part;
  CUtensorMap ma, mb;
  if (!make_map_bf16(&ma, X, M, K, ldx, 64, 64) || !make_map_bf16(&mb, dY, M, N, ldy, 64, 64)) {
    set_error("tt_linear_wgrad_bf16: cuTensorMapEncodeTiled failed");
    return TT_ERR_CUDA;
  }
  int rc = launch_gemm(ma, mb, p, eff, s);
  if (rc != TT_OK) return rc;
  if (eff > 1) gt_splitk_reduce_kernel<<<ceil_div((int64_t)N * K, 256), 256, 0, s>>>(part, (int64_t)N * K, eff, dW, (int64_t)N * K);
  if (dbias) {
    double* cpart = reinterpret_cast<double*>((char*)workspace + (((int64_t)splits * N * K * 4 + 255) / 256 * 256));
    const int chunks = (int)((M + GCS_ROWS - 1) / GCS_ROWS);
    colsum_bf16_partial_kernel<<<dim3(ceil_div(N, 32), chunks), 256, 0, s>>>((const __nv_bfloat16*)dY, ldy, M, N, cpart);
    colsum_bf16_final_kernel<<<ceil_div(N, 8), 256, 0, s>>>(cpart, chunks, N, dbias);
  }
  TT_LAUNCH_CHECK("linear_wgrad_bf16");
  return TT_OK;
}

}  // extern "C"
