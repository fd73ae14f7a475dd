// Fused tower kernels of the bf16 train step (tcgen05 / TMEM / TMA). The tower MLP of the reference is
//     Linear -> ReLU -> BatchNorm1d -> Dropout -> Linear -> F.normalize          (base_tower.py:87-97,142-145)
// and at the reference's sizes every stage is a few microseconds of launch + set-up latency, so the chain is cut into the
// fewest kernels the BatchNorm batch statistics allow (a grid-wide dependency before and after them):
//   forward   [Linear1 + ReLU + per-tile column statistics]  (gemm_tc.cu, tt_linear_forward_bf16_stats)
//             [BN finalize + apply + dropout -> Linear2 + bias -> L2 normalise -> bf16 (+ pre-scaled copy)]   tower_tail_fwd_kernel
//   backward  [L2-normalise backward -> bf16 -> dgrad Linear2 -> dropout mask -> BN-backward column sums]     tower_head_bwd_kernel
//             [BN backward apply + ReLU mask -> bf16 -> dgrad Linear1 (N tiles of 128 through TMEM) + bias-gradient sums]
//                                                                                                             tower_mid_bwd_kernel
// A CTA owns 128 rows (= TMEM lanes). The elementwise stage runs on four warps and writes its result as the K-major,
// 128B-swizzled A operand straight into shared memory (and to HBM where a weight gradient needs it later); one more warp loads
// the weight tile(s) by TMA and issues tcgen05.mma; the same four warps read the accumulator back (thread == row, so row
// reductions — the L2 norm — need no shuffles). Column reductions (BatchNorm sums, bias gradients) go through shared memory in a
// fixed order: everything is deterministic. fp32 per-tile partials are folded in fp64 by the consumer.
#include "tc_common.cuh"

namespace tt {

constexpr int TF_ROWS = 128;
constexpr int TF_THREADS = 160;      // warps 0-3 elementwise + epilogue, warp 4 TMA + MMA issue + TMEM allocation
constexpr int TF_MID_THREADS = 192;  // mid backward: warp 4 MMA, warp 5 TMA producer

__device__ __forceinline__ void bar128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// byte offset of 16-byte chunk `c` (8 bf16) of row `r` inside a K-major 128B-swizzled operand made of 64-element atoms
__device__ __forceinline__ uint32_t swz_off(int r, int c) { return (uint32_t)((c >> 3) * ATOM_BYTES + r * 128 + (((c & 7) ^ (r & 7)) << 4)); }

// keep * 1/(1-p) for 8 consecutive elements starting at flat index idx0 (a multiple of 8): the same Philox stream as
// dropout_keep() in common.cuh, two calls instead of eight
__device__ __forceinline__ void keep8(const uint8_t* mask_in, const tt_adam_hyper* hyp, uint32_t stream, int64_t idx0, float p,
                                      float inv_keep, float (&k)[8]) {
  if (p <= 0.f) {
#pragma unroll
    for (int u = 0; u < 8; ++u) k[u] = 1.f;
    return;
  }
  if (mask_in) {
    const uint2 mv = *reinterpret_cast<const uint2*>(mask_in + idx0);
    const uint8_t* mb = reinterpret_cast<const uint8_t*>(&mv);
#pragma unroll
    for (int u = 0; u < 8; ++u) k[u] = mb[u] ? inv_keep : 0.f;
    return;
  }
  const uint64_t seed = hyp->rng_seed, step = hyp->rng_step;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const uint64_t idx = (uint64_t)idx0 + 4 * g;
    const uint4 c = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), (uint32_t)step, stream ^ (uint32_t)(step >> 32) * 0x9E3779B9u);
    const uint4 r = philox4x32_10(c, key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) k[4 * g + u] = ((float)(w[u] >> 8) * (1.0f / 16777216.0f) >= p) ? inv_keep : 0.f;
  }
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 t = __bfloat1622float2(h[u]);
    f[2 * u] = t.x;
    f[2 * u + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int u = 0; u < 4; ++u) h[u] = __floats2bfloat162_rn(f[2 * u], f[2 * u + 1]);
  return o;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// column sums over the 128 rows of a bf16 [128 x N] K-major swizzled operand tile in shared memory (N = 64 or 128), fixed order:
// thread t -> column t % N, row range (t / N) * (128 * N / 128)...; `red` is 128 floats of scratch. Result for column n in out[n]
// (written by threads < N). All 128 elementwise threads call it.
__device__ __forceinline__ void tile_colsum_bf16(const unsigned char* tile, int N, float* red, float* out_global, int tid) {
  const int n = tid % N, q = tid / N;     // N = 64: q in {0,1}; N = 128: q = 0
  const int rows_per = N == 64 ? 64 : 128;
  float s = 0.f;
  const int c = n >> 3, e = n & 7;
  for (int r = q * rows_per; r < (q + 1) * rows_per; ++r)
    s += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + swz_off(r, c) + e * 2));
  red[tid] = s;
  bar128();
  if (tid < N) out_global[tid] = N == 64 ? red[tid] + red[tid + 64] : red[tid];
  bar128();
}


// Fold per-tile partials part[t][C] (fp32, C % 4 == 0, C <= 512) over n_tiles tiles into out[C] (fp64, shared memory) with all 128
// elementwise threads: thread -> one float4 column group and every (128 / groups)-th tile, eight independent 16-byte loads in
// flight, then the tile subsets are added in a fixed order — deterministic, and one L2 round trip instead of n_tiles of them.
// scratch: 512 doubles of shared memory. Ends with a barrier (out is valid for every thread).
__device__ __forceinline__ void fold_partials(const float* __restrict__ part, int n_tiles, int C, double* out, double* scratch, int tid) {
  const int groups = C >> 2;
  const int n_sub = 128 / groups > 0 ? 128 / groups : 1;
  if (tid < groups * n_sub) {
    const int g = tid % groups, q = tid / groups;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int t = q;
    for (; t + 7 * n_sub < n_tiles; t += 8 * n_sub) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(part + (int64_t)(t + u * n_sub) * C) + g);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0 += (double)v[u].x; a1 += (double)v[u].y; a2 += (double)v[u].z; a3 += (double)v[u].w;
      }
    }
    for (; t < n_tiles; t += n_sub) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + (int64_t)t * C) + g);
      a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
    }
    double* dst = scratch + (q * groups + g) * 4;
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
  }
  bar128();
  for (int c = tid; c < C; c += 128) {
    double t = 0.0;
    for (int q = 0; q < n_sub; ++q) t += scratch[q * C + c];
    out[c] = t;
  }
  bar128();
}

// ============================================================================================== forward tail
struct TailFwdParams {
  const __nv_bfloat16* H;  // [M, H1] ReLU(Linear1) as stored by the producing GEMM
  int64_t M;
  int H1, F;
  const float* stat_part;  // [n_tiles][2][H1] per-tile column sums of h and h^2 (single GPU)
  int n_tiles;
  const double* sums;      // or [2][H1] global sums (SyncBN: reduced + all-reduced by the caller); count = rows they cover
  long long count;
  const float *gamma, *beta;
  float eps, momentum;
  float *mean, *rstd, *rm, *rv;
  long long* nbt;
  float p_drop;
  const uint8_t* mask_in;
  const tt_adam_hyper* hyp;
  uint32_t rng_stream;
  __nv_bfloat16* Y;        // [M, H1] out: dropout(bn(h)) — the input of Linear2, kept for its weight gradient
  const float* bias2;
  __nv_bfloat16 *E, *Es;   // [M, F] out: unit rows, and unit rows * es_scale (Es may be NULL)
  float es_scale;
  float* nrm;              // [M] out
};

struct FusedSmem {
  uint64_t w_full, a_full, acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(TF_THREADS, 1)
tower_tail_fwd_kernel(const __grid_constant__ CUtensorMap map_w2, const TailFwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NA = p.H1 >> 6;                 // K atoms
  unsigned char* sA = base;                 // NA x 16 KB
  unsigned char* sB = sA + NA * ATOM_BYTES; // NA x (F x 128 B)
  double* s_sum = reinterpret_cast<double*>(sB + NA * p.F * 128);  // [2 * H1] folded statistics, then 512 doubles of scratch
  float* sc = reinterpret_cast<float*>(s_sum + 2 * p.H1 + 512);
  float* sh = sc + p.H1;
  FusedSmem* sb = reinterpret_cast<FusedSmem*>(sh + p.H1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TF_ROWS;
  const uint32_t tmem_cols = p.F <= 64 ? 64 : 128;
  if (tid == 0) {
    mbar_init(&sb->w_full, 1);
    mbar_init(&sb->a_full, 128);
    mbar_init(&sb->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sb->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  pdl_wait();  // programmatic dependent launch: the set-up above overlapped the predecessor's tail
  pdl_launch_dependents();

  if (warp == 4) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&sb->w_full, (uint32_t)(NA * p.F * 128));
      for (int a = 0; a < NA; ++a) tma_load_2d(sB + a * p.F * 128, &map_w2, a * 64, 0, &sb->w_full);  // box {64 k, F n}
    }
    __syncwarp();
    mbar_wait(&sb->w_full, 0);
    mbar_wait(&sb->a_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, p.F, 0, 0);
    const uint64_t dA = make_desc(smem_u32(sA), 16, 1024), dB = make_desc(smem_u32(sB), 16, 1024);
    if (elect_one_sync()) {
      for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, dA + (uint64_t)(a * (ATOM_BYTES >> 4) + 2 * k), dB + (uint64_t)(a * ((p.F * 128) >> 4) + 2 * k), idesc,
                    (a > 0 || k > 0) ? 1u : 0u);
      umma_commit(&sb->acc_full);
    }
    __syncwarp();
  } else {
    // this thread's first eight 16-byte chunks of h are requested before anything else: their latency hides behind the
    // statistics fold (chunk idx = tid + 128 it: consecutive threads -> consecutive chunks, coalesced)
    const int cpr = p.H1 >> 3;  // chunks per row == iterations per thread (8 or 16)
    uint4 hv[8];
    auto load_h = [&](int it, uint4& v) {
      const int idx = tid + it * 128;
      const int r = idx / cpr, c = idx - r * cpr;
      const int64_t mr = m0 + r;
      v = make_uint4(0u, 0u, 0u, 0u);
      if (mr < p.M) v = __ldg(reinterpret_cast<const uint4*>(p.H + mr * p.H1) + c);
    };
#pragma unroll
    for (int u = 0; u < 8; ++u) load_h(u, hv[u]);
    // ---- BatchNorm statistics -> per-column scale / shift (every CTA folds the per-tile partials itself)
    if (!p.sums) fold_partials(p.stat_part, p.n_tiles, 2 * p.H1, s_sum, s_sum + 2 * p.H1, tid);
    for (int n = tid; n < p.H1; n += 128) {
      const double s0 = p.sums ? p.sums[n] : s_sum[n];
      const double s1 = p.sums ? p.sums[p.H1 + n] : s_sum[p.H1 + n];
      const double mu = s0 / (double)p.count;
      double var = s1 / (double)p.count - mu * mu;  // biased
      if (var < 0.0) var = 0.0;
      const float rs = (float)(1.0 / sqrt(var + (double)p.eps));
      const float g = p.gamma[n];
      sc[n] = g * rs;
      sh[n] = p.beta[n] - (float)mu * g * rs;
      if (blockIdx.x == 0) {
        p.mean[n] = (float)mu;
        p.rstd[n] = rs;
        if (p.rm) p.rm[n] = (1.f - p.momentum) * p.rm[n] + p.momentum * (float)mu;
        if (p.rv) {
          const double unb = p.count > 1 ? var * ((double)p.count / (double)(p.count - 1)) : var;
          p.rv[n] = (1.f - p.momentum) * p.rv[n] + p.momentum * (float)unb;
        }
      }
    }
    if (blockIdx.x == 0 && tid == 0 && p.nbt) *p.nbt += 1;
    bar128();
    // ---- y = dropout(bn(h)) -> bf16 -> HBM (for the weight gradient) and the swizzled A operand
    const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    for (int base_it = 0; base_it < cpr; base_it += 8) {
      if (base_it > 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) load_h(base_it + u, hv[u]);
      }
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8) {
        const int idx = tid + (base_it + u8) * 128;
        const int r = idx / cpr, c = idx - r * cpr;
        const int64_t mr = m0 + r;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (mr < p.M) {
          float f[8], k[8];
          unpack8(hv[u8], f);
          keep8(p.mask_in, p.hyp, p.rng_stream, mr * p.H1 + c * 8, p.p_drop, inv_keep, k);
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = fmaf(f[u], sc[c * 8 + u], sh[c * 8 + u]) * k[u];
          o = pack8(f);
          *(reinterpret_cast<uint4*>(p.Y + mr * p.H1) + c) = o;
        }
        *reinterpret_cast<uint4*>(sA + swz_off(r, c)) = o;
      }
    }
    fence_proxy_async();
    mbar_arrive(&sb->a_full);
    // ---- epilogue: z = acc + bias -> L2 norm -> unit rows (bf16), thread == row
    const int64_t m = m0 + tid;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    mbar_wait(&sb->acc_full, 0);
    tc_fence_after();
    float ss = 0.f;
    for (int c32 = 0; c32 < p.F / 32; ++c32) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_addr + c32 * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float z = __uint_as_float(r[j]) + __ldg(p.bias2 + c32 * 32 + j);
        ss = fmaf(z, z, ss);
      }
    }
    const float nr = fmaxf(sqrtf(ss), 1e-12f);
    if (m < p.M) p.nrm[m] = nr;
    for (int c32 = 0; c32 < p.F / 32; ++c32) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_addr + c32 * 32, r);
      tmem_ld_wait();
      if (m < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float e[8], es[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float z = __uint_as_float(r[q * 8 + u]) + __ldg(p.bias2 + c32 * 32 + q * 8 + u);
            e[u] = z / nr;
            es[u] = e[u] * p.es_scale;
          }
          *(reinterpret_cast<uint4*>(p.E + m * p.F + c32 * 32) + q) = pack8(e);
          if (p.Es) *(reinterpret_cast<uint4*>(p.Es + m * p.F + c32 * 32) + q) = pack8(es);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, tmem_cols);
  }
}

// ============================================================================================== backward head
struct HeadBwdParams {
  const float* DE;          // [M, F] gradient of the unit rows
  const __nv_bfloat16* E;   // [M, F]
  const float* nrm;         // [M]
  int64_t M;
  int F, H1;
  __nv_bfloat16* DZ16;      // [M, F] out: gradient of Linear2's output (its weight gradient reads it)
  float* dz_colpart;        // [n_tiles][F] out: per-tile column sums of dz (bias gradient of Linear2)
  const __nv_bfloat16* H;   // [M, H1]
  const float *mean, *rstd;
  float p_drop;
  const uint8_t* mask_in;
  const tt_adam_hyper* hyp;
  uint32_t rng_stream;
  __nv_bfloat16* DYP;       // [M, H1] out: dy * keep / (1 - p): gradient w.r.t. the BatchNorm output
  float* bn_part;           // [n_tiles][2][H1] out: column sums of dyp and dyp * xhat
};

__global__ void __launch_bounds__(TF_THREADS, 1)
tower_head_bwd_kernel(const __grid_constant__ CUtensorMap map_w2, const HeadBwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NA = p.F >> 6;                        // K atoms of dz (reduction over F)
  const int NB = p.H1 >> 6;                       // 64-column MN blocks of W2
  unsigned char* sA = base;                       // NA x 16 KB
  unsigned char* sB = sA + NA * ATOM_BYTES;       // NB blocks of [F rows x 128 B]
  float* s_mean = reinterpret_cast<float*>(sB + NB * p.F * 128);
  float* s_rstd = s_mean + p.H1;
  float* red = s_rstd + p.H1;                     // 2 x 128 x 33 + 512 floats
  FusedSmem* sb = reinterpret_cast<FusedSmem*>(red + 2 * 128 * 33 + 512);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TF_ROWS;
  const uint32_t tmem_cols = p.H1 <= 64 ? 64 : 128;
  if (tid == 0) {
    mbar_init(&sb->w_full, 1);
    mbar_init(&sb->a_full, 128);
    mbar_init(&sb->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sb->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  pdl_wait();  // programmatic dependent launch: the set-up above overlapped the predecessor's tail
  pdl_launch_dependents();

  if (warp == 4) {
    // dy[m, k] = sum_n dz[m, n] W2[n, k]: B is MN-major (k contiguous), the F reduction rows of a 64-column block back to back
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&sb->w_full, (uint32_t)(NB * p.F * 128));
      for (int b = 0; b < NB; ++b) tma_load_2d(sB + b * p.F * 128, &map_w2, b * 64, 0, &sb->w_full);  // box {64 k, F n}
    }
    __syncwarp();
    mbar_wait(&sb->w_full, 0);
    mbar_wait(&sb->a_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, p.H1, 0, 1);
    const uint64_t dA = make_desc(smem_u32(sA), 16, 1024), dB = make_desc(smem_u32(sB), (uint32_t)(p.F * 128), 1024);
    if (elect_one_sync()) {
      for (int kk = 0; kk < p.F / 16; ++kk)
        umma_bf16(tmem, dA + (uint64_t)((kk >> 2) * (ATOM_BYTES >> 4) + 2 * (kk & 3)), dB + (uint64_t)(kk * (2048 >> 4)), idesc,
                  kk > 0 ? 1u : 0u);
      umma_commit(&sb->acc_full);
    }
    __syncwarp();
  } else {
    for (int n = tid; n < p.H1; n += 128) {
      s_mean[n] = p.mean[n];
      s_rstd[n] = p.rstd[n];
    }
    // ---- dz = (de - e (e . de)) / |z|: F / 8 lanes per row (one 16-byte chunk of e, two of de each: coalesced), the row dot
    // through F / 8-lane shuffles; warp w owns rows 32 w .. 32 w + 31 of the tile
    const int64_t m = m0 + tid;  // the epilogue's row of this thread
    {
      const int lpr = p.F >> 3;                 // lanes per row: 8 or 16
      const int rpi = 32 / lpr;                 // rows per warp iteration: 4 or 2
      const int lane = tid & 31, c = lane % lpr;
      for (int base_it = 0; base_it < 32 / rpi; base_it += 8) {
        uint4 ev[8];
        float4 dv0[8], dv1[8];
        float invs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // all loads of eight iterations first
          const int64_t mr = m0 + warp * 32 + (base_it + u) * rpi + lane / lpr;
          ev[u] = make_uint4(0u, 0u, 0u, 0u);
          dv0[u] = dv1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          invs[u] = 0.f;
          if (mr < p.M) {
            ev[u] = __ldg(reinterpret_cast<const uint4*>(p.E + mr * p.F) + c);
            dv0[u] = __ldg(reinterpret_cast<const float4*>(p.DE + mr * p.F) + 2 * c);
            dv1[u] = __ldg(reinterpret_cast<const float4*>(p.DE + mr * p.F) + 2 * c + 1);
            invs[u] = 1.f / __ldg(p.nrm + mr);
          }
        }
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) {
          const int r = warp * 32 + (base_it + u8) * rpi + lane / lpr;
          const int64_t mr = m0 + r;
          float e[8], dot = 0.f;
          unpack8(ev[u8], e);
          const float d[8] = {dv0[u8].x, dv0[u8].y, dv0[u8].z, dv0[u8].w, dv1[u8].x, dv1[u8].y, dv1[u8].z, dv1[u8].w};
#pragma unroll
          for (int u = 0; u < 8; ++u) dot = fmaf(e[u], d[u], dot);
          for (int o = lpr >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);  // same value on the row's lanes
          float dz[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) dz[u] = (d[u] - e[u] * dot) * invs[u8];
          const uint4 o = pack8(dz);  // rows past M: zeros
          if (mr < p.M) *(reinterpret_cast<uint4*>(p.DZ16 + mr * p.F) + c) = o;
          *reinterpret_cast<uint4*>(sA + swz_off(r, c)) = o;
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(&sb->a_full);
    bar128();  // the whole dz tile is in shared memory (also orders s_mean / s_rstd)
    tile_colsum_bf16(sA, p.F, red, p.dz_colpart + (int64_t)blockIdx.x * p.F, tid);
    // ---- epilogue: dyp = dy * keep; column sums of dyp and dyp * xhat
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    mbar_wait(&sb->acc_full, 0);
    tc_fence_after();
    const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    float* red0 = red;
    float* red1 = red + 128 * 33;
    float* red2 = red + 2 * 128 * 33;
    for (int c32 = 0; c32 < p.H1 / 32; ++c32) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_addr + c32 * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float a[8], b[8];
        if (m < p.M) {
          float h[8], k[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.H + m * p.H1 + c32 * 32) + q), h);
          keep8(p.mask_in, p.hyp, p.rng_stream, m * p.H1 + c32 * 32 + q * 8, p.p_drop, inv_keep, k);
#pragma unroll
          for (int u = 0; u < 8; ++u) a[u] = __uint_as_float(r[q * 8 + u]) * k[u];
          const uint4 o = pack8(a);
          *(reinterpret_cast<uint4*>(p.DYP + m * p.H1 + c32 * 32) + q) = o;
          unpack8(o, a);  // the sums are over what the next kernel reads: the bf16-rounded values
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int n = c32 * 32 + q * 8 + u;
            b[u] = a[u] * ((h[u] - s_mean[n]) * s_rstd[n]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) a[u] = b[u] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          red0[tid * 33 + q * 8 + u] = a[u];
          red1[tid * 33 + q * 8 + u] = b[u];
        }
      }
      bar128();
      {
        const int col = tid & 31, qq = tid >> 5;
        float s0 = 0.f, s1 = 0.f;
        for (int rr = qq * 32; rr < qq * 32 + 32; ++rr) {
          s0 += red0[rr * 33 + col];
          s1 += red1[rr * 33 + col];
        }
        red2[qq * 32 + col] = s0;
        red2[128 + qq * 32 + col] = s1;
      }
      bar128();
      if (tid < 64) {
        const int which = tid >> 5, col = tid & 31;
        const float* src = red2 + which * 128 + col;
        p.bn_part[((int64_t)blockIdx.x * 2 + which) * p.H1 + c32 * 32 + col] = (src[0] + src[32]) + (src[64] + src[96]);
      }
      bar128();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, tmem_cols);
  }
}

// ============================================================================================== backward middle
constexpr int MID_WSTAGES = 3;  // weight-tile ring
constexpr int MID_ABUF = 2;     // accumulator buffers in TMEM (2 x 128 columns: two CTAs fit one SM's 512 columns)

struct MidBwdParams {
  const __nv_bfloat16* DYP;  // [M, H1]
  const __nv_bfloat16* H;    // [M, H1]
  int64_t M;
  int H1, LD;                // LD = width of Linear1's input (= the concat buffer) = number of output columns here
  const float* bn_part;      // [n_tiles][2][H1] (single GPU) ...
  int n_tiles;
  const double* sums;        // ... or [2][H1] global sums (SyncBN)
  long long count;
  const float *mean, *rstd, *gamma;
  float *dgamma, *dbeta;     // block 0 publishes affine_scale * sums
  float affine_scale;
  const float* dz_colpart;   // [n_tiles][F]: block 0 folds Linear2's bias gradient
  int F;
  float* dbias2;
  __nv_bfloat16* DH;         // [M, H1] out: gradient of Linear1's output (after the ReLU mask)
  float* dh_colpart;         // [n_tiles][H1] out
  __nv_bfloat16* DC;         // [M, LD] out: gradient of the concat buffer
  int H0;                    // columns [0, H0) are the dense projection's output: their column sums are its bias gradient
  float* dc_colpart;         // [n_tiles][H0] out
};

struct MidSmem {
  uint64_t a_full;
  uint64_t w_full[MID_WSTAGES], w_empty[MID_WSTAGES];
  uint64_t acc_full[MID_ABUF], acc_empty[MID_ABUF];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(TF_MID_THREADS, 1)
tower_mid_bwd_kernel(const __grid_constant__ CUtensorMap map_w1, const MidBwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NA = p.H1 >> 6;
  const int wtile_bytes = p.H1 * 256;              // [H1 reduction rows] x [128 output columns] bf16 = 2 MN blocks of H1 x 128 B
  unsigned char* sA = base;                        // NA x 16 KB: dh, K-major
  unsigned char* sW = sA + NA * ATOM_BYTES;        // MID_WSTAGES weight tiles
  unsigned char* sStage = sW + MID_WSTAGES * wtile_bytes;  // 4 warps x [32 rows x 256 B] output staging
  double* s_sum = reinterpret_cast<double*>(sStage + 4 * 8192);  // [2 * H1] folded sums, then 512 doubles of scratch
  float* s_col = reinterpret_cast<float*>(s_sum + 2 * p.H1 + 512);  // mu, rs, g, a0, a1: 5 x H1
  float* red = s_col + 5 * p.H1;                   // 4 x 128 floats
  MidSmem* sb = reinterpret_cast<MidSmem*>(red + 512);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TF_ROWS;
  // blockIdx.y = column group: the 128-column output tiles [t_begin, t_end) of this CTA (every group recomputes the cheap dh
  // tile; group 0 stores it). More CTAs than row tiles: the epilogue — the bulk of the kernel — spreads over the machine.
  const int n_out_all = (p.LD + 127) / 128;
  const int t_begin = (int)((long long)blockIdx.y * n_out_all / gridDim.y);
  const int t_end = (int)((long long)(blockIdx.y + 1) * n_out_all / gridDim.y);
  const bool lead = blockIdx.y == 0;
  if (tid == 0) {
    mbar_init(&sb->a_full, 128);
    for (int i = 0; i < MID_WSTAGES; ++i) {
      mbar_init(&sb->w_full[i], 1);
      mbar_init(&sb->w_empty[i], 1);
    }
    for (int i = 0; i < MID_ABUF; ++i) {
      mbar_init(&sb->acc_full[i], 1);
      mbar_init(&sb->acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sb->tmem_base, MID_ABUF * 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  pdl_wait();  // programmatic dependent launch: the set-up above overlapped the predecessor's tail
  pdl_launch_dependents();

  if (warp == 5) {
    // ---- weight tiles: W1[:, k0 : k0 + 128] as two MN-major blocks of [H1 rows x 64 columns] (out-of-range columns read as zero)
    int st = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&sb->w_empty[st], ph ^ 1);
      if (elect_one_sync()) {
        unsigned char* dst = sW + st * wtile_bytes;
        mbar_arrive_expect_tx(&sb->w_full[st], (uint32_t)wtile_bytes);
        tma_load_2d(dst, &map_w1, t * 128, 0, &sb->w_full[st]);                       // box {64 k, H1 n}
        tma_load_2d(dst + p.H1 * 128, &map_w1, t * 128 + 64, 0, &sb->w_full[st]);
      }
      __syncwarp();
      if (++st == MID_WSTAGES) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 4) {
    const uint32_t idesc = make_idesc(128, 128, 0, 1);
    const uint64_t dA = make_desc(smem_u32(sA), 16, 1024);
    const uint64_t dW0 = make_desc(smem_u32(sW), (uint32_t)(p.H1 * 128), 1024);
    mbar_wait(&sb->a_full, 0);
    tc_fence_after();
    int st = 0, ab = 0;
    uint32_t ph = 0, aph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&sb->w_full[st], ph);
      mbar_wait(&sb->acc_empty[ab], aph ^ 1);
      tc_fence_after();
      const uint64_t dW = dW0 + (uint64_t)(st * (wtile_bytes >> 4));
      if (elect_one_sync()) {
        for (int kk = 0; kk < p.H1 / 16; ++kk)
          umma_bf16(tmem + ab * 128, dA + (uint64_t)((kk >> 2) * (ATOM_BYTES >> 4) + 2 * (kk & 3)), dW + (uint64_t)(kk * (2048 >> 4)), idesc,
                    kk > 0 ? 1u : 0u);
        umma_commit(&sb->w_empty[st]);
        umma_commit(&sb->acc_full[ab]);
      }
      __syncwarp();
      if (++st == MID_WSTAGES) {
        st = 0;
        ph ^= 1;
      }
      if (++ab == MID_ABUF) {
        ab = 0;
        aph ^= 1;
      }
    }
  } else {
    float* s_mu = s_col;
    float* s_rs = s_col + p.H1;
    float* s_g = s_col + 2 * p.H1;
    float* s_a0 = s_col + 3 * p.H1;
    float* s_a1 = s_col + 4 * p.H1;
    // ---- BatchNorm-backward sums -> per-column coefficients; block 0 publishes the affine and bias gradients
    const int cpr = p.H1 >> 3;
    uint4 hv[8], dv[8];
    auto load_hd = [&](int it, uint4& vh, uint4& vd) {
      const int idx = tid + it * 128;
      const int r = idx / cpr, c = idx - r * cpr;
      const int64_t mr = m0 + r;
      vh = vd = make_uint4(0u, 0u, 0u, 0u);
      if (mr < p.M) {
        vh = __ldg(reinterpret_cast<const uint4*>(p.H + mr * p.H1) + c);
        vd = __ldg(reinterpret_cast<const uint4*>(p.DYP + mr * p.H1) + c);
      }
    };
#pragma unroll
    for (int u = 0; u < 8; ++u) load_hd(u, hv[u], dv[u]);  // in flight underneath the folds below
    if (blockIdx.x == 0 && lead && p.dbias2) {  // bias gradient of Linear2: fold of the head kernel's per-tile column sums of dz
      fold_partials(p.dz_colpart, p.n_tiles, p.F, s_sum, s_sum + 2 * p.H1, tid);
      for (int f = tid; f < p.F; f += 128) p.dbias2[f] = (float)s_sum[f];
      bar128();
    }
    if (!p.sums) fold_partials(p.bn_part, p.n_tiles, 2 * p.H1, s_sum, s_sum + 2 * p.H1, tid);
    for (int n = tid; n < p.H1; n += 128) {
      const double S0 = p.sums ? p.sums[n] : s_sum[n];
      const double S1 = p.sums ? p.sums[p.H1 + n] : s_sum[p.H1 + n];
      const float rs = p.rstd[n];
      s_mu[n] = p.mean[n];
      s_rs[n] = rs;
      s_g[n] = rs * p.gamma[n];
      s_a0[n] = (float)S0 * (1.f / (float)p.count);
      s_a1[n] = (float)S1 * (1.f / (float)p.count);
      if (blockIdx.x == 0 && lead) {
        if (p.dbeta) p.dbeta[n] = p.affine_scale * (float)S0;
        if (p.dgamma) p.dgamma[n] = p.affine_scale * (float)S1;
      }
    }
    bar128();
    // ---- dh = relu'(h) rstd gamma (dyp - mean(dyp) - xhat mean(dyp xhat)) -> bf16 -> HBM + swizzled A operand
    for (int base_it = 0; base_it < cpr; base_it += 8) {
      if (base_it > 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) load_hd(base_it + u, hv[u], dv[u]);
      }
#pragma unroll
      for (int u8 = 0; u8 < 8; ++u8) {
        const int idx = tid + (base_it + u8) * 128;
        const int r = idx / cpr, c = idx - r * cpr;
        const int64_t mr = m0 + r;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (mr < p.M) {
          float h[8], d[8];
          unpack8(hv[u8], h);
          unpack8(dv[u8], d);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int n = c * 8 + u;
            const float xhat = (h[u] - s_mu[n]) * s_rs[n];
            const float dh = s_g[n] * (d[u] - s_a0[n] - xhat * s_a1[n]);
            d[u] = h[u] > 0.f ? dh : 0.f;
          }
          o = pack8(d);
          if (lead) *(reinterpret_cast<uint4*>(p.DH + mr * p.H1) + c) = o;
        }
        *reinterpret_cast<uint4*>(sA + swz_off(r, c)) = o;
      }
    }
    fence_proxy_async();
    mbar_arrive(&sb->a_full);
    bar128();
    if (lead) tile_colsum_bf16(sA, p.H1, red, p.dh_colpart + (int64_t)blockIdx.x * p.H1, tid);
    // ---- epilogue: per 128-column output tile, TMEM -> bf16 -> per-warp staging (XOR-swizzled 16-byte chunks) -> coalesced rows
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    unsigned char* stg = sStage + warp * 8192;
    int ab = 0;
    uint32_t aph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&sb->acc_full[ab], aph);
      tc_fence_after();
#pragma unroll
      for (int c32 = 0; c32 < 4; ++c32) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_addr + ab * 128 + c32 * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(r[q * 8 + u]);
          const int chunk = c32 * 4 + q;  // 16 chunks of 16 bytes per 256-byte row
          *reinterpret_cast<uint4*>(stg + lane * 256 + ((chunk ^ (lane & 15)) << 4)) = pack8(f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->acc_empty[ab]);
      const int k0 = t * 128;
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const int rr = it * 2 + (lane >> 4), chunk = lane & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 256 + ((chunk ^ (rr & 15)) << 4));
        const int64_t m = m0 + warp * 32 + rr;
        if (m < p.M && k0 + chunk * 8 < p.LD) *reinterpret_cast<uint4*>(p.DC + m * p.LD + k0 + chunk * 8) = v;
      }
      if (k0 < p.H0) {
        // bias gradient of the dense projection: column sums of this tile over the CTA's rows (rows past M hold zeros: their dh
        // rows are zero). lane -> columns 4*lane .. 4*lane+3 of the warp's 32 staged rows, then the four warps in order.
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int rr = 0; rr < 32; ++rr) {
          const int chunk = lane >> 1;
          const uint2 v = *reinterpret_cast<const uint2*>(stg + rr * 256 + ((chunk ^ (rr & 15)) << 4) + (lane & 1) * 8);
          const float4 f = unpack_bf16x4(v);
          s4[0] += f.x; s4[1] += f.y; s4[2] += f.z; s4[3] += f.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) red[warp * 128 + lane * 4 + u] = s4[u];
        bar128();
        if (k0 + tid < p.H0) p.dc_colpart[(int64_t)blockIdx.x * p.H0 + k0 + tid] = (red[tid] + red[128 + tid]) + (red[256 + tid] + red[384 + tid]);
        bar128();
      }
      __syncwarp();
      if (++ab == MID_ABUF) {
        ab = 0;
        aph ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, MID_ABUF * 128);
  }
}

// ---------------------------------------------------------------- small reducers
// sums[c] = sum over tiles of part[t][c] (fp64, tile order): the SyncBN hand-off (the caller all-reduces `sums`)
__global__ void __launch_bounds__(256) colpart_reduce_f64_kernel(const float* __restrict__ part, int n_tiles, int C, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int t = 0; t < n_tiles; ++t) s += (double)part[(int64_t)t * C + c];
  sums[c] = s;
}
// up to four bias-gradient segments in one launch: out[c] = sum over tiles of part[t][c]
struct ColSegs {
  const float* part[4];
  float* out[4];
  int C[4];
  int n_tiles;
};
__global__ void __launch_bounds__(256) colpart_reduce_multi_kernel(const ColSegs s) {
  const int seg = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= s.C[seg]) return;
  double a = 0.0;
  const float* src = s.part[seg] + c;
  const int64_t ld = s.C[seg];
  int t = 0;
  for (; t + 16 <= s.n_tiles; t += 16) {  // sixteen partials in flight, added in tile order
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = src[(int64_t)(t + u) * ld];
#pragma unroll
    for (int u = 0; u < 16; ++u) a += (double)v[u];
  }
  for (; t < s.n_tiles; ++t) a += (double)src[(int64_t)t * ld];
  s.out[seg][c] = (float)a;
}

static bool fused_dims_ok(int H1, int F) { return (H1 == 64 || H1 == 128) && (F == 64 || F == 128); }

}  // namespace tt

using namespace tt;

extern "C" {

int tt_tower_fused_supported(int32_t H0, int32_t H1, int32_t F, int32_t LD) {
  return fused_dims_ok(H1, F) && H0 > 0 && (H0 % 128) == 0 && (LD % 8) == 0 && LD >= H0 ? 1 : 0;
}

int tt_tower_tail_forward(const void* h, int64_t M, int32_t H1, int32_t F, const float* stat_part, int32_t n_tiles, const double* sums,
                          int64_t count, const float* gamma, const float* beta, float eps, float momentum, float* mean, float* rstd,
                          float* running_mean, float* running_var, int64_t* num_batches_tracked, float dropout_p,
                          const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream, void* y, const void* W2_bf16,
                          const float* bias2, void* e, void* e_scaled, float e_scale, float* nrm, tt_stream_t stream) {
  TT_CHECK_ARG(h && gamma && beta && mean && rstd && y && W2_bf16 && bias2 && e && nrm && M > 0 && count > 0, "tt_tower_tail_forward: bad arguments");
  TT_CHECK_ARG(fused_dims_ok(H1, F), "tt_tower_tail_forward: hidden width and final dim must be 64 or 128 (got %d, %d)", H1, F);
  TT_CHECK_ARG(sums || (stat_part && n_tiles == (M + 127) / 128), "tt_tower_tail_forward: statistics missing");
  TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || mask_in || d_hyper), "tt_tower_tail_forward: bad dropout arguments");
  TailFwdParams p = {};
  p.H = (const __nv_bfloat16*)h; p.M = M; p.H1 = H1; p.F = F; p.stat_part = stat_part; p.n_tiles = n_tiles; p.sums = sums;
  p.count = count; p.gamma = gamma; p.beta = beta; p.eps = eps; p.momentum = momentum; p.mean = mean; p.rstd = rstd;
  p.rm = running_mean; p.rv = running_var; p.nbt = (long long*)num_batches_tracked; p.p_drop = dropout_p; p.mask_in = mask_in;
  p.hyp = d_hyper; p.rng_stream = rng_stream; p.Y = (__nv_bfloat16*)y; p.bias2 = bias2; p.E = (__nv_bfloat16*)e;
  p.Es = (__nv_bfloat16*)e_scaled; p.es_scale = e_scale; p.nrm = nrm;
  CUtensorMap mw;
  if (!make_map_bf16(&mw, W2_bf16, F, H1, H1, 64, F)) {
    set_error("tt_tower_tail_forward: cuTensorMapEncodeTiled failed");
    return TT_ERR_CUDA;
  }
  const int NA = H1 / 64;
  const size_t smem = 1024 + (size_t)NA * ATOM_BYTES + (size_t)NA * F * 128 + (2 * H1 + 512) * 8 + 2 * H1 * 4 + sizeof(FusedSmem) + 64;
  TT_CUDA(cudaFuncSetAttribute(tower_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TT_CUDA(launch_pdl(tower_tail_fwd_kernel, dim3(ceil_div(M, TF_ROWS)), dim3(TF_THREADS), smem, (cudaStream_t)stream, mw, p));
  TT_LAUNCH_CHECK("tower_tail_fwd_kernel");
  return TT_OK;
}

int tt_tower_head_backward(const float* de, const void* e, const float* nrm, int64_t M, int32_t F, int32_t H1, const void* W2_bf16,
                           void* dz_bf16, float* dz_colpart, const void* h, const float* mean, const float* rstd, float dropout_p,
                           const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream, void* dyp, float* bn_part,
                           tt_stream_t stream) {
  TT_CHECK_ARG(de && e && nrm && W2_bf16 && dz_bf16 && dz_colpart && h && mean && rstd && dyp && bn_part && M > 0, "tt_tower_head_backward: bad arguments");
  TT_CHECK_ARG(fused_dims_ok(H1, F), "tt_tower_head_backward: hidden width and final dim must be 64 or 128 (got %d, %d)", H1, F);
  TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || mask_in || d_hyper), "tt_tower_head_backward: bad dropout arguments");
  HeadBwdParams p = {};
  p.DE = de; p.E = (const __nv_bfloat16*)e; p.nrm = nrm; p.M = M; p.F = F; p.H1 = H1; p.DZ16 = (__nv_bfloat16*)dz_bf16;
  p.dz_colpart = dz_colpart; p.H = (const __nv_bfloat16*)h; p.mean = mean; p.rstd = rstd; p.p_drop = dropout_p; p.mask_in = mask_in;
  p.hyp = d_hyper; p.rng_stream = rng_stream; p.DYP = (__nv_bfloat16*)dyp; p.bn_part = bn_part;
  CUtensorMap mw;
  if (!make_map_bf16(&mw, W2_bf16, F, H1, H1, 64, F)) {
    set_error("tt_tower_head_backward: cuTensorMapEncodeTiled failed");
    return TT_ERR_CUDA;
  }
  const size_t smem = 1024 + (size_t)(F / 64) * ATOM_BYTES + (size_t)(H1 / 64) * F * 128 + 2 * H1 * 4 + (2 * 128 * 33 + 512) * 4 +
                      sizeof(FusedSmem) + 64;
  TT_CUDA(cudaFuncSetAttribute(tower_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TT_CUDA(launch_pdl(tower_head_bwd_kernel, dim3(ceil_div(M, TF_ROWS)), dim3(TF_THREADS), smem, (cudaStream_t)stream, mw, p));
  TT_LAUNCH_CHECK("tower_head_bwd_kernel");
  return TT_OK;
}

int tt_tower_mid_backward(const void* dyp, const void* h, int64_t M, int32_t H1, int32_t LD, const void* W1_bf16, const float* bn_part,
                          int32_t n_tiles, const double* sums, int64_t count, const float* mean, const float* rstd, const float* gamma,
                          float* dgamma, float* dbeta, float affine_grad_scale, const float* dz_colpart, int32_t F, float* dbias2,
                          void* dh, float* dh_colpart, void* dc, int32_t H0, float* dc_colpart, tt_stream_t stream) {
  TT_CHECK_ARG(dyp && h && W1_bf16 && mean && rstd && gamma && dh && dh_colpart && dc && dc_colpart && M > 0 && count > 0,
               "tt_tower_mid_backward: bad arguments");
  TT_CHECK_ARG(tt_tower_fused_supported(H0, H1, F > 0 ? F : 64, LD), "tt_tower_mid_backward: unsupported dims (H0 %d, H1 %d, LD %d)", H0, H1, LD);
  TT_CHECK_ARG(n_tiles == (M + 127) / 128 && (sums || bn_part) && (!dbias2 || dz_colpart), "tt_tower_mid_backward: partials missing");
  MidBwdParams p = {};
  p.DYP = (const __nv_bfloat16*)dyp; p.H = (const __nv_bfloat16*)h; p.M = M; p.H1 = H1; p.LD = LD; p.bn_part = bn_part;
  p.n_tiles = n_tiles; p.sums = sums; p.count = count; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.dgamma = dgamma;
  p.dbeta = dbeta; p.affine_scale = affine_grad_scale; p.dz_colpart = dz_colpart; p.F = F; p.dbias2 = dbias2;
  p.DH = (__nv_bfloat16*)dh; p.dh_colpart = dh_colpart; p.DC = (__nv_bfloat16*)dc; p.H0 = H0; p.dc_colpart = dc_colpart;
  CUtensorMap mw;
  if (!make_map_bf16(&mw, W1_bf16, H1, LD, LD, 64, H1)) {
    set_error("tt_tower_mid_backward: cuTensorMapEncodeTiled failed");
    return TT_ERR_CUDA;
  }
  const size_t smem = 1024 + (size_t)(H1 / 64) * ATOM_BYTES + (size_t)MID_WSTAGES * H1 * 256 + 4 * 8192 + (2 * H1 + 512) * 8 +
                      5 * H1 * 4 + 512 * 4 + sizeof(MidSmem) + 64;
  TT_CUDA(cudaFuncSetAttribute(tower_mid_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // column groups: as many as keep roughly two CTAs per SM busy (two fit: ~100 KB of shared memory, 256 TMEM columns each)
  const int row_tiles = ceil_div(M, TF_ROWS), n_out = (LD + 127) / 128;
  int groups = (2 * sm_count()) / row_tiles;
  if (groups < 1) groups = 1;
  if (groups > n_out) groups = n_out;
  TT_CUDA(launch_pdl(tower_mid_bwd_kernel, dim3(row_tiles, groups), dim3(TF_MID_THREADS), smem, (cudaStream_t)stream, mw, p));
  TT_LAUNCH_CHECK("tower_mid_bwd_kernel");
  return TT_OK;
}

int tt_colpart_reduce_f64(const float* part, int32_t n_tiles, int32_t C, double* sums, tt_stream_t stream) {
  TT_CHECK_ARG(part && sums && n_tiles > 0 && C > 0, "tt_colpart_reduce_f64: bad arguments");
  colpart_reduce_f64_kernel<<<ceil_div(C, 256), 256, 0, (cudaStream_t)stream>>>(part, n_tiles, C, sums);
  TT_LAUNCH_CHECK("colpart_reduce_f64_kernel");
  return TT_OK;
}

int tt_colpart_reduce2(const float* part0, int32_t C0, float* out0, const float* part1, int32_t C1, float* out1, int32_t n_tiles,
                       tt_stream_t stream) {
  TT_CHECK_ARG(part0 && out0 && C0 > 0 && n_tiles > 0 && (C1 == 0 || (part1 && out1)), "tt_colpart_reduce2: bad arguments");
  ColSegs s = {};
  s.part[0] = part0; s.out[0] = out0; s.C[0] = C0;
  s.part[1] = part1; s.out[1] = out1; s.C[1] = C1;
  s.n_tiles = n_tiles;
  const int cmax = C0 > C1 ? C0 : C1;
  colpart_reduce_multi_kernel<<<dim3(ceil_div(cmax, 256), C1 > 0 ? 2 : 1), 256, 0, (cudaStream_t)stream>>>(s);
  TT_LAUNCH_CHECK("colpart_reduce_multi_kernel");
  return TT_OK;
}

}  // extern "C"
