// Similarity logits + symmetric softmax-CE on the 5th-gen tensor cores (tcgen05) — bf16 path.
//
// One CTA owns a 128-row block of X and walks a chunk of 128-column tiles of Y:
//     S_tile[128x128] = X_blk[128xF] · Y_tile[128xF]^T      tcgen05.mma, fp32 accumulator in TMEM
// The logits never leave the SM: the epilogue warps pull the tile out of TMEM with tcgen05.ld (one row per thread, so row
// reductions need no shuffles) and take E = exp2 of the shifted logits. The epilogue is MUFU- and issue-bound, so it does
// the bare minimum per logit and everything else is moved to the tensor core or to the small reduce kernels:
//   forward : per-row running sum(E) and maximum; the arg-max is kept per 16-column chunk and resolved afterwards, sum_j S_ij
//             is X_i . (sum_j Y_j);
//   backward: dS_ij = E_ij (ix_i + iy_j) - target_ij. Only E is formed per logit; it is parked as bf16 in shared memory
//             (K-major, 128B-swizzled) and fed straight back to the tensor core against the operand [Y | Ys], Ys = diag(iy) Y:
//             [E.Y | E.Ys][128 x 2F] accumulates in TMEM across the chunk, the read-out applies ix_i, and the target terms are
//             a rank-one and a diagonal correction added by tc_dx_reduce_kernel.
// Column statistics / the other tower's gradient come from running the same kernel on the transposed problem
// (X and Y swapped), so no cross-CTA column reduction and no atomics: everything is deterministic.
//
// Warp roles (608 threads): warps 0-15 epilogue (TMEM lane quarter == warp % 4, 32-column slice == warp / 4), warp 16 issues
// S = X Y^T and owns the TMEM allocation, warp 17 issues the backward MMA, warp 18 loads operands (TMA: one elected lane
// issues cp.async.bulk.tensor with 128B swizzle; the generic-proxy fallback loads 16-byte chunks and applies the same swizzle
// by hand). The issuing warps run warp-uniform code with elect.sync around the tcgen05 / TMA instructions, so their operands
// stay in uniform registers.
// Pipelines: Y (and Ys) tiles through a 3-4 stage smem ring (full/empty mbarriers), S through 3 (forward) or 2 (backward)
// TMEM buffers read one 16-column half ahead, E through 2 smem buffers, the backward accumulator double-buffered in TMEM
// when it fits; tcgen05.commit arrives on the barriers when the MMAs that read a buffer have retired.
#include "tc_common.cuh"
#include <float.h>

namespace tt {

constexpr int TCM = 128;          // rows per CTA block (UMMA M)
constexpr int TCN = 128;          // columns per tile  (UMMA N)
constexpr int TC_STAGES = 6;      // Y smem ring (upper bound; the depth in use is TcParams::n_stages)
constexpr int TC_SBUF = 3;        // forward: S accumulators in TMEM (3 x 128 columns), the MMA runs two tiles ahead of the
                                  // epilogue; backward: 2, the other 256 columns hold the [E.Y | E.Ys] accumulator(s)
constexpr int TC_EPI_WARPS = 16;  // four per SM sub-partition (the epilogue is MUFU/latency bound): rows = TMEM lane quarter
                                  // (warp % 4), columns = the 32-column slice warp / 4 of every 128-column tile
constexpr int TC_MMA_WARP = 16;   // issues S = X Y^T (and owns the TMEM allocation)
constexpr int TC_MMA2_WARP = 17;  // backward: issues dX += dS Y, so neither issue chain waits for the other
constexpr int TC_LOAD_WARP = 18;
constexpr int TC_THREADS = 608;   // 16 epilogue + 2 mma + 1 loader warps
constexpr int TC_LOADER_THREADS = 32;

struct TcParams {
  const __nv_bfloat16* X;
  const __nv_bfloat16* Y;
  int64_t Bx, By;
  int F;                 // 64 or 128
  int n_col_tiles;       // ceil(By / 128)
  long long total_tiles; // row blocks * column tiles: CTA c owns the contiguous tile range [c*T/G, (c+1)*T/G) of the row-block-
                         // major tile sequence, i.e. one or two long items instead of many short ones
  float c1, c0;          // e = exp2(s_raw * c1 + c0):  c1 = inv_temp*log2e, c0 = -inv_temp*log2e (unused when the X operand is
                         // pre-scaled by c1: then the accumulator IS the exponent and e = exp2(s) is unshifted)
  float inv_temp;        // diag = raw * inv_temp (pre-scaled X: ln 2, which maps the exponent back to S / T)
  int64_t diag_offset;
  int use_tma;
  int n_stages;          // Y smem ring depth (<= TC_STAGES)
  int n_xbuf;            // X row-block buffers (1 or 2)
  int n_dbuf;            // backward: [E.Y | E.Ys] accumulators in TMEM
  int n_sbuf;            // S buffers in TMEM: 3 forward; backward 2 (F = 64) or 1 (F = 128), see the TMEM map in the kernel
  int poly;              // pre-scaled passes: 1 = every fourth exp2 runs as a polynomial on the FMA pipe (tt_debug_simce_knobs)
  long long* dbg_ts;     // optional timeline buffer (block 0): [4 roles][512 tiles][4 events] clock64 stamps, see scripts/time_simce.py
  // forward outputs: partials per (chunk, row)
  float* row_e;
  float* row_max;
  int32_t* row_arg;
  float* diag;
  int want_metrics;
  // backward
  const __nv_bfloat16* Ys;  // [By][F] = bf16(gscale * inv_y[j] * Y[j,:]) (column scales folded into the second MMA operand)
  const float* inv_x;    // [Bx] (unscaled)
  float gscale;
  float* dx_part;        // [max parts][Bx][F]: part = owning CTA - first CTA of the row block. Deferred row scale (MODE 2):
                         // [max parts][Bx][2F], the raw halves [E.Y | E.Ys]; the row sums of E go to row_e like in the forward
};

// MODE of simce_tc_kernel: the forward statistics pass, the gradient pass with a known row scale, and the gradient pass that
// also accumulates this row block's sum of E (the row scale is only known once the pass is over: the reduce kernel applies it)
enum { TC_FWD = 0, TC_BWD = 1, TC_BWD_SUMS = 2 };

struct TcSmem {  // carved from dynamic smem after the 1024-aligned tiles
  uint64_t x_full[2], x_empty[2];  // the X row block is double-buffered so the next work item's block loads early
  uint64_t y_full[TC_STAGES], y_empty[TC_STAGES];
  uint64_t s_full[TC_SBUF], s_empty[TC_SBUF];
  uint64_t p_full[2], p_empty[2];
  uint64_t d_full[2], d_empty[2];
  uint32_t tmem_base;
};

// copy a [128 x F] bf16 row-major tile (rows row0.. of G, zero past `rows`) into 128B-swizzled K-major atoms
__device__ __forceinline__ void manual_tile_load(unsigned char* dst, const __nv_bfloat16* G, int64_t row0, int64_t rows, int F,
                                                 int lt /* 0..63 */) {
  const int chunks_per_row = F >> 3;            // 16-byte chunks per row
  const int total = TCM * chunks_per_row;       // 1024 or 2048
  for (int base = 0; base < total; base += TC_LOADER_THREADS * 8) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * TC_LOADER_THREADS + lt;
      const int r = idx / chunks_per_row, c = idx - r * chunks_per_row;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < total && row0 + r < rows) v[u] = __ldg(reinterpret_cast<const uint4*>(G + (row0 + r) * F) + c);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * TC_LOADER_THREADS + lt;
      if (idx < total) {
        const int r = idx / chunks_per_row, c = idx - r * chunks_per_row;
        const int atom = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(dst + atom * ATOM_BYTES + r * 128 + ((cc ^ (r & 7)) << 4)) = v[u];
      }
    }
  }
}

// ---- epilogue helpers -------------------------------------------------------------------------------
// Experiment knob: every N-th logit takes exp2 on the FMA pipe (ex2_poly) instead of the MUFU pipe. Measured at B = 8192: N = 4
// and N = 8 are both SLOWER (forward 26.3 -> 28.2 / 27.2 us): with four warps per scheduler the epilogue is bound by instruction
// issue and dependency latency before it is bound by the 16/clk/SM MUFU rate, so trading one MUFU for eight FMA-pipe
// instructions loses. Left at 0 (off).
constexpr int TC_POLY_EVERY = 0;
__device__ __forceinline__ float tc_exp2(float x, int j) {
  return (TC_POLY_EVERY > 0 && (j % (TC_POLY_EVERY > 0 ? TC_POLY_EVERY : 1)) == TC_POLY_EVERY - 1) ? ex2_poly(x) : ex2f(x);
}
// Pre-scaled passes (no FFMA, no shift): the epilogue is bound by the MUFU pipe alone (16 ex2/clk/SM: 1024 cycles per 128x128
// tile), with issue slots to spare. scripts/microbench/mufu2.cu: MUFU alone 15.9 exps/clk/SM, the degree-4 polynomial alone
// 12.0, three MUFU + one polynomial 20.4 — so every fourth exponential of a chunk goes to the FMA pipe when POLY is set.
template <bool POLY>
__device__ __forceinline__ float tc_exp2p(float x, int j) {
  return (POLY && (j & 3) == 3) ? ex2_poly(x) : ex2f(x);
}
constexpr int TC_CW = 16;  // columns per TMEM load / epilogue chunk: half of a warp's 32-column slice

// The epilogue is issue-bound, not just MUFU-bound (four warps per scheduler): every instruction per logit counts.
// Forward, per logit: FFMA (scale), MUFU.EX2, FADD (row sum-exp) and half an FMNMX3 (row maximum). The arg-max is tracked per
// 16-column chunk only (three instructions per chunk); tc_rows_reduce_kernel resolves the column inside the winning chunk.
// sum_j S_ij is not accumulated here at all: it equals X_i . (sum_j Y_j) and is formed in the reduce kernel.
// GENERIC handles ragged column tiles and the diagonal.
// PRE: the X operand was pre-scaled by c1 = log2(e) / T when it was written (tt_l2norm_forward_scaled), so the accumulator is the
// exponent itself: no FFMA per logit, and E = exp2(s) is unshifted (|s| <= 57.8 at the smallest temperature: no overflow in
// fp32 / bf16 even summed over 2^16 columns).
template <bool GENERIC, bool PRE, bool POLY>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&r)[TC_CW], int cbase, int By, float c1, float c0, int dcol, float (&acc)[4],
                                          float& best, int32_t& best_cb, float& dval, bool& have_diag) {
  float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
  for (int j = 0; j < TC_CW; ++j) {
    const float raw = __uint_as_float(r[j]);
    if (GENERIC) {
      const int col = cbase + j;
      if (col < By) {
        acc[j & 3] += PRE ? tc_exp2p<POLY>(raw, j) : tc_exp2(fmaf(raw, c1, c0), j);
        m4[j & 3] = fmaxf(m4[j & 3], raw);
        if (col == dcol) {
          dval = raw;
          have_diag = true;
        }
      }
    } else {
      acc[j & 3] += PRE ? tc_exp2p<POLY>(raw, j) : tc_exp2(fmaf(raw, c1, c0), j);
      m4[j & 3] = fmaxf(m4[j & 3], raw);
    }
  }
  const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  if (m > best) {  // strict: the first chunk that attains the maximum wins (chunks are visited in ascending column order)
    best = m;
    best_cb = cbase;
  }
}

// Backward: E = exp(S - 1/T) for TC_CW logits -> bf16 pairs, stored straight back into tensor memory (tcgen05.st) where the
// second MMA reads them as its A operand: the E tile never touches shared memory. Nothing else happens per logit: with dS_ij = E_ij (ix_i + iy_j) - target_ij the row scale ix_i is applied to the
// accumulator, the column scale iy_j is folded into the second MMA operand (Ys = diag(iy) Y), and the target terms are rank-one
// / diagonal corrections added by tc_dx_reduce_kernel:
//   dX_i = ix_i (E Y)_i + (E Ys)_i - t_off sum_j Y_j - (t_diag - t_off) Y_dcol(i).
// Ragged tiles need no masks: rows / columns past the edge are zero-filled operands (they contribute nothing, or are not stored).
template <bool PRE, bool SUMS, bool POLY>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&r)[TC_CW], float c1, float c0, uint32_t p_taddr, float (&acc)[4]) {
  uint32_t w[TC_CW / 2];
#pragma unroll
  for (int u = 0; u < TC_CW / 2; ++u) {
    const float e0 = PRE ? tc_exp2p<POLY>(__uint_as_float(r[2 * u]), 2 * u) : tc_exp2(fmaf(__uint_as_float(r[2 * u]), c1, c0), 2 * u);
    const float e1 = PRE ? tc_exp2p<POLY>(__uint_as_float(r[2 * u + 1]), 2 * u + 1)
                         : tc_exp2(fmaf(__uint_as_float(r[2 * u + 1]), c1, c0), 2 * u + 1);
    if (SUMS) {  // this row's sum of E rides along (one FADD per logit): the transposed gradient pass IS the column-statistics pass
      acc[(2 * u) & 3] += e0;
      acc[(2 * u + 1) & 3] += e1;
    }
    __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);  // low half = even column
    w[u] = *reinterpret_cast<uint32_t*>(&t);
  }
  tmem_st8(p_taddr, w);
}

// ---- work partition: the T = row blocks x column tiles tiles, row-block-major, cut into G equal contiguous ranges ----
__host__ __device__ __forceinline__ long long tc_range_begin(int c, long long T, int G) { return (long long)c * T / G; }
// the CTA whose range holds tile g: the largest c with floor(c T / G) <= g, i.e. c T < (g + 1) G — one division (the reduce kernels
// evaluate this per thread, in front of their loads)
__host__ __device__ __forceinline__ int tc_cta_of_tile(long long g, long long T, int G) {
  return (int)((unsigned long long)((g + 1) * G - 1) / (unsigned long long)T);
}
// the pieces of row block rb are numbered by owning CTA, starting at the CTA that owns the block's first tile
__host__ __device__ __forceinline__ int tc_parts_of_row_block(int rb, int n_ct, long long T, int G) {
  return tc_cta_of_tile((long long)rb * n_ct + n_ct - 1, T, G) - tc_cta_of_tile((long long)rb * n_ct, T, G) + 1;
}

template <int MODE, bool PRE>
__global__ void __launch_bounds__(TC_THREADS, 1)
simce_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                const __grid_constant__ CUtensorMap map_ys, const TcParams p) {
  // TMEM map (512 columns). Forward: three S buffers. Backward: n_sbuf S buffers | [E.Y | E.Ys] accumulator (2F columns) |
  // two E buffers (64 columns each: 128 bf16 per row, two per column): F = 64 -> 256 + 128 + 128, F = 128 -> 128 + 256 + 128.
  constexpr bool BWD = MODE != TC_FWD;
  constexpr bool SUMS = MODE == TC_BWD_SUMS;
  const int NSB = p.n_sbuf;
  extern __shared__ unsigned char smem_raw[];
  if (p.dbg_ts && threadIdx.x == 0) {  // per-CTA wall-clock stamps (ns): entry, after set-up, exit
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg_ts[2 * 2048 + 1024 + blockIdx.x * 4 + 0] = (long long)g;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NA = p.F >> 6;                      // swizzle atoms along K (1 or 2)
  const int tile_bytes = NA * ATOM_BYTES;       // X / Y tile
  const int stage_bytes = (BWD ? 2 : 1) * tile_bytes;  // backward stages hold the Y tile and, right behind it, the Ys tile
  const int NST = p.n_stages;
  unsigned char* sX = base;                     // n_xbuf tiles (double-buffered across work items when they fit)
  unsigned char* sY = sX + p.n_xbuf * tile_bytes;  // NST stages
  TcSmem* sb = reinterpret_cast<TcSmem*>(sY + NST * stage_bytes);

  // warp index through a shuffle: the compiler then treats the role dispatch (and everything computed under it from
  // kernel parameters) as warp-uniform, which keeps tcgen05 operands in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_loader_arrivals = p.use_tma ? 1 : TC_LOADER_THREADS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->x_full[i], n_loader_arrivals);
      mbar_init(&sb->x_empty[i], 1);
    }
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&sb->y_full[i], n_loader_arrivals);
      mbar_init(&sb->y_empty[i], 1);
    }
    for (int i = 0; i < TC_SBUF; ++i) {
      mbar_init(&sb->s_full[i], 1);
      mbar_init(&sb->s_empty[i], TC_EPI_WARPS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->p_full[i], TC_EPI_WARPS);
      mbar_init(&sb->p_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->d_full[i], 1);
      mbar_init(&sb->d_empty[i], 8);  // the eight read-out warps
    }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(&sb->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  // programmatic dependent launch: everything above (barriers, tensor memory) overlapped the predecessor's tail; its results are
  // read from here on. The successor may be scheduled as soon as every CTA of this grid has got this far.
  pdl_wait();
  pdl_launch_dependents();
  if (p.dbg_ts && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg_ts[2 * 2048 + 1024 + blockIdx.x * 4 + 1] = (long long)g;
  }
  const uint32_t tmem_d = tmem + NSB * TCN;  // backward accumulator(s): n_dbuf x 2F columns behind the S buffers
  const int n_dbuf = p.n_dbuf;
  const uint32_t tmem_p = tmem_d + n_dbuf * 2 * p.F;  // two E buffers of 64 columns
  const long long g_begin = tc_range_begin((int)blockIdx.x, p.total_tiles, (int)gridDim.x);
  const long long g_end = tc_range_begin((int)blockIdx.x + 1, p.total_tiles, (int)gridDim.x);

  if (warp == TC_LOAD_WARP) {
    // =========================== operand loader ===========================
    const int lt = lane;
    uint32_t x_iter = 0, y_iter = 0;
    int st = 0;           // Y ring cursor kept incrementally: no runtime division on the per-tile path
    uint32_t ph = 0;
    for (long long g = g_begin; g < g_end; ++x_iter) {
      const int rb = (int)(g / p.n_col_tiles), t0 = (int)(g - (long long)rb * p.n_col_tiles);
      const int t1 = t0 + (int)min((long long)(p.n_col_tiles - t0), g_end - g);
      g += t1 - t0;
      const int xb = p.n_xbuf == 2 ? (x_iter & 1) : 0;
      unsigned char* sXb = sX + xb * tile_bytes;
      mbar_wait(&sb->x_empty[xb], ((p.n_xbuf == 2 ? (x_iter >> 1) : x_iter) & 1) ^ 1);
      if (p.use_tma) {
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&sb->x_full[xb], (uint32_t)tile_bytes);
          for (int a = 0; a < NA; ++a) tma_load_2d(sXb + a * ATOM_BYTES, &map_x, a * 64, rb * TCM, &sb->x_full[xb]);
        }
      } else {
        manual_tile_load(sXb, p.X, (int64_t)rb * TCM, p.Bx, p.F, lt);
        fence_proxy_async();
        mbar_arrive(&sb->x_full[xb]);
      }
      for (int t = t0; t < t1; ++t, ++y_iter) {
        if (p.dbg_ts && blockIdx.x == 0 && lt == 0 && y_iter < 500) p.dbg_ts[2 * 2048 + y_iter * 4 + 0] = clock64();
        mbar_wait(&sb->y_empty[st], ph ^ 1);
        if (p.dbg_ts && blockIdx.x == 0 && lt == 0 && y_iter < 500) p.dbg_ts[2 * 2048 + y_iter * 4 + 1] = clock64();
        unsigned char* dst = sY + st * stage_bytes;
        if (p.use_tma) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&sb->y_full[st], (uint32_t)stage_bytes);
            for (int a = 0; a < NA; ++a) tma_load_2d(dst + a * ATOM_BYTES, &map_y, a * 64, t * TCN, &sb->y_full[st]);
            if (BWD)
              for (int a = 0; a < NA; ++a) tma_load_2d(dst + tile_bytes + a * ATOM_BYTES, &map_ys, a * 64, t * TCN, &sb->y_full[st]);
          }
        } else {
          manual_tile_load(dst, p.Y, (int64_t)t * TCN, p.By, p.F, lt);
          if (BWD) manual_tile_load(dst + tile_bytes, p.Ys, (int64_t)t * TCN, p.By, p.F, lt);
          fence_proxy_async();
          mbar_arrive(&sb->y_full[st]);
        }
        if (++st == NST) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // =========================== MMA issuer 1: S = X Y^T ===========================
    // The whole warp runs the loop (uniform control flow, operands in uniform registers); one elected lane issues.
    const uint32_t idesc1 = make_idesc(TCM, TCN, 0, 0);
    const int ksteps1 = p.F >> 4;
    // descriptors differ only in their 14-bit start-address field: build them once, then one 64-bit add per operand
    const uint64_t dX0 = make_desc(smem_u32(sX), 16, 1024), dY0 = make_desc(smem_u32(sY), 16, 1024);
    const uint32_t tile16 = (uint32_t)tile_bytes >> 4, stage16 = (uint32_t)stage_bytes >> 4;
    uint32_t x_iter = 0, s_iter = 0;
    int y_st = 0, s_buf = 0;  // ring cursors kept incrementally: this warp's per-tile latency chain paces the CTA
    uint32_t y_ph = 0, s_ph = 0;
    for (long long g = g_begin; g < g_end; ++x_iter) {
      const int t0 = (int)(g % p.n_col_tiles);
      const int nt = (int)min((long long)(p.n_col_tiles - t0), g_end - g);
      g += nt;
      const int xb = p.n_xbuf == 2 ? (x_iter & 1) : 0;
      const uint64_t dXb = dX0 + (uint64_t)(xb * tile16);
      mbar_wait(&sb->x_full[xb], (p.n_xbuf == 2 ? (x_iter >> 1) : x_iter) & 1);
      for (int t = 0; t < nt; ++t, ++s_iter) {
        const bool ts = p.dbg_ts && blockIdx.x == 0 && lane == 0 && s_iter < 500;
        if (ts) p.dbg_ts[0 * 2048 + s_iter * 4 + 0] = clock64();
        mbar_wait(&sb->y_full[y_st], y_ph);
        if (ts) p.dbg_ts[0 * 2048 + s_iter * 4 + 1] = clock64();
        mbar_wait(&sb->s_empty[s_buf], s_ph ^ 1);
        if (ts) p.dbg_ts[0 * 2048 + s_iter * 4 + 2] = clock64();
        tc_fence_after();
        const uint64_t dYs = dY0 + (uint64_t)(y_st * stage16);
        const uint32_t d_tmem = tmem + s_buf * TCN;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, dXb + 2 * k, dYs + 2 * k, idesc1, k > 0 ? 1u : 0u);
          if (ksteps1 > 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, dXb + (ATOM_BYTES >> 4) + 2 * k, dYs + (ATOM_BYTES >> 4) + 2 * k, idesc1, 1u);
          }
          umma_commit(&sb->s_full[s_buf]);
          if (!BWD) umma_commit(&sb->y_empty[y_st]);          // backward: the stage is released after MMA2 has read it too
          if (t == nt - 1) umma_commit(&sb->x_empty[xb]);     // every MMA that reads this item's X has been issued
        }
        __syncwarp();
        if (ts) p.dbg_ts[0 * 2048 + s_iter * 4 + 3] = clock64();
        if (++y_st == NST) {
          y_st = 0;
          y_ph ^= 1;
        }
        if (++s_buf == NSB) {
          s_buf = 0;
          s_ph ^= 1;
        }
      }
    }
  } else if (warp == TC_MMA2_WARP) {
    // ====================== MMA issuer 2 (backward): [E.Y | E.Ys] += E(K-major) * [Y | Ys](MN-major) ======================
    // The stage holds the Y tile and the Ys tile back to back, i.e. 2F/64 consecutive 64-column atoms: one N = 2F operand.
    if (BWD) {
      const uint32_t idesc2 = make_idesc(TCM, 2 * p.F, 0, 1);
      const uint64_t dYmn0 = make_desc(smem_u32(sY), ATOM_BYTES, 1024);
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      uint32_t p_iter = 0;
      int y_st = 0, d_buf = 0;
      uint32_t y_ph = 0, d_ph = 0;
      for (long long g = g_begin; g < g_end;) {
        const int t0 = (int)(g % p.n_col_tiles);
        const int nt = (int)min((long long)(p.n_col_tiles - t0), g_end - g);
        g += nt;
        const uint32_t d_tmem = tmem_d + d_buf * 2 * p.F;
        for (int t = 0; t < nt; ++t, ++p_iter) {
          const int pb = p_iter & 1;
          const bool ts2 = p.dbg_ts && blockIdx.x == 0 && lane == 0 && p_iter < 500;
          if (ts2) p.dbg_ts[3 * 2048 + p_iter * 4 + 0] = clock64();
          mbar_wait(&sb->p_full[pb], (p_iter >> 1) & 1);
          if (ts2) p.dbg_ts[3 * 2048 + p_iter * 4 + 1] = clock64();
          mbar_wait(&sb->y_full[y_st], y_ph);  // long complete (MMA1 consumed it); acquires the TMA-written stage for this warp
          if (t == 0) mbar_wait(&sb->d_empty[d_buf], d_ph ^ 1);
          tc_fence_after();
          const uint32_t a_tmem = tmem_p + pb * (TCN / 2);
          const uint64_t dYm = dYmn0 + (uint64_t)(y_st * stage16);
          if (elect_one_sync()) {
#pragma unroll
            for (int kk = 0; kk < TCN / 16; ++kk)  // K = 16 logits columns per step = 8 TMEM columns of E
              umma_bf16_ts(d_tmem, a_tmem + kk * 8, dYm + (uint64_t)(kk * (2048 >> 4)), idesc2, (t > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&sb->p_empty[pb]);
            umma_commit(&sb->y_empty[y_st]);
            if (t == nt - 1) umma_commit(&sb->d_full[d_buf]);
          }
          __syncwarp();
          if (ts2) p.dbg_ts[3 * 2048 + p_iter * 4 + 2] = clock64();
          if (++y_st == NST) {
            y_st = 0;
            y_ph ^= 1;
          }
        }
        if (++d_buf == n_dbuf) {
          d_buf = 0;
          d_ph ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    // thread == one row of the block (TMEM lane quarter = warp % 4) and one 32-column slice of every tile (warp / 4)
    const int quarter = warp & 3, slice = warp >> 2;
    const int r_in_blk = quarter * 32 + lane;
    const uint32_t t_base = tmem + ((uint32_t)(quarter * 32) << 16) + slice * 32;  // this thread's lanes / columns of S buffer 0
    const int By = (int)p.By;
    const float c1 = p.c1, c0 = p.c0;
    const bool poly = p.poly != 0;
    const int t_part = (By & (TCN - 1)) ? p.n_col_tiles - 1 : -1;  // the ragged last column tile, if any
    uint32_t p_iter = 0, e_iter = 0;
    // backward: the accumulator of an item is read out one tile into the NEXT item, so the epilogue never idles while the last
    // E tile of an item drains through MMA2 (with a single accumulator the next item's first MMA2 waits for that read-out, the
    // epilogue itself does not wait for anything)
    int d_buf = 0;
    uint32_t d_ph = 0;
    bool pend = false;
    float* pend_out = nullptr;
    int pend_buf = 0;
    uint32_t pend_ph = 0;
    bool pend_ok = false;
    float pend_ixs = 0.f;
    int s_cur = 0;           // S ring cursor of the tile being processed; the phase flips when it wraps
    uint32_t s_cur_ph = 0;
    // read one [E.Y | E.Ys] accumulator out of TMEM and apply the row scale: out = ix * (E.Y) + (E.Ys). Two of the four warps
    // of each scheduler do this (half of the F columns of their 32 rows each): they need no MUFU meanwhile, so the other two
    // run ahead into the next tile (one S / E buffer of slack) and the read-out stays off the MUFU-bound critical path.
    auto drain_d = [&](float* out, int db, uint32_t ph, bool ok, float ix, uint32_t (&r)[TC_CW]) {
      if (slice >= 2) return;
      mbar_wait(&sb->d_full[db], ph);
      tc_fence_after();
      const int h0 = slice * (p.F >> 1), h1 = h0 + (p.F >> 1);
      const uint32_t ta = tmem_d + db * 2 * p.F + ((uint32_t)(quarter * 32) << 16);
      for (int h = h0; h < h1; h += TC_CW) {
        uint32_t r2[TC_CW];
        tmem_ld16(ta + h, r);
        tmem_ld16(ta + p.F + h, r2);
        tmem_ld_wait_dep(r);
        tmem_ld_wait_dep(r2);
        if (ok) {
          if (SUMS) {  // the row scale is not known yet: both halves leave as they are ([E.Y | E.Ys], 2F floats per row and part)
#pragma unroll
            for (int q = 0; q < TC_CW / 4; ++q) {
              *reinterpret_cast<uint4*>(out + h + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
              *reinterpret_cast<uint4*>(out + p.F + h + q * 4) = make_uint4(r2[q * 4], r2[q * 4 + 1], r2[q * 4 + 2], r2[q * 4 + 3]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < TC_CW / 4; ++q)
              *reinterpret_cast<float4*>(out + h + q * 4) =
                  make_float4(fmaf(ix, __uint_as_float(r[q * 4]), __uint_as_float(r2[q * 4])),
                              fmaf(ix, __uint_as_float(r[q * 4 + 1]), __uint_as_float(r2[q * 4 + 1])),
                              fmaf(ix, __uint_as_float(r[q * 4 + 2]), __uint_as_float(r2[q * 4 + 2])),
                              fmaf(ix, __uint_as_float(r[q * 4 + 3]), __uint_as_float(r2[q * 4 + 3])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb->d_empty[db]);
    };
    int it_iter = 0;
    for (long long g = g_begin; g < g_end; ++it_iter) {
      const bool istamp = p.dbg_ts && blockIdx.x == 0 && threadIdx.x == 0 && it_iter < 64;
      if (istamp) p.dbg_ts[1 * 2048 + 1024 + it_iter * 4 + 0] = clock64();
      const int rb = (int)(g / p.n_col_tiles), t0 = (int)(g - (long long)rb * p.n_col_tiles);
      const int t1 = t0 + (int)min((long long)(p.n_col_tiles - t0), g_end - g);
      g += t1 - t0;
      const bool last_item = g >= g_end;
      const int ch = (int)blockIdx.x - tc_cta_of_tile((long long)rb * p.n_col_tiles, p.total_tiles, (int)gridDim.x);  // part index
      const int64_t row = (int64_t)rb * TCM + r_in_blk;
      const bool row_ok = row < p.Bx;
      // this row's positive column, and the (at most two) column tiles the block's diagonal crosses
      const int64_t d0 = p.diag_offset + (int64_t)rb * TCM;
      const int dcol = (d0 + r_in_blk >= 0 && d0 + r_in_blk < p.By) ? (int)(d0 + r_in_blk) : -1;
      const int t_d0 = (d0 + TCM > 0 && d0 < p.By) ? (int)((d0 < 0 ? 0 : d0) / TCN) : -1;
      const int t_d1 = (d0 + TCM > 0 && d0 < p.By) ? (int)((d0 + TCM - 1) / TCN) : -1;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      float best = -FLT_MAX, dval = 0.f;
      int32_t best_cb = 0x7fffffff;
      bool have_diag = false;
      float ixs = 0.f;
      if (BWD && !SUMS) ixs = row_ok ? p.gscale * p.inv_x[row] : 0.f;
      // Each warp reads its 32-column slice of an S tile as two 16-column TMEM loads, software-pipelined one half ahead:
      // the tcgen05.ld of the next half (or of the next tile's first half) is in flight while the current half is processed
      // out of the other register set, and the S buffer returns to the MMA warp as soon as its second half sits in registers.
      uint32_t rA[TC_CW], rB[TC_CW];
      const bool stamp0 = p.dbg_ts && blockIdx.x == 0 && threadIdx.x == 0;
      mbar_wait(&sb->s_full[s_cur], s_cur_ph);
      tc_fence_after();
      tmem_ld16(t_base + s_cur * TCN, rA);
      if (istamp) p.dbg_ts[1 * 2048 + 1024 + it_iter * 4 + 1] = clock64();
      for (int t = t0; t < t1; ++t) {
        const bool generic = !BWD && (t == t_d0 || t == t_d1 || t == t_part);
        const int cb = t * TCN + slice * 32;
        int s_nxt = s_cur + 1;
        uint32_t s_nxt_ph = s_cur_ph;
        if (s_nxt == NSB) {
          s_nxt = 0;
          s_nxt_ph ^= 1;
        }
        const bool stamp = stamp0 && e_iter < 500;
        if (stamp) p.dbg_ts[1 * 2048 + e_iter * 4 + 0] = clock64();
        tmem_ld_wait_dep(rA);                        // first half landed
        tmem_ld16(t_base + s_cur * TCN + TC_CW, rB);  // second half in flight ...
        int pb = 0;
        uint32_t p_taddr = 0;
        if (BWD) {  // a free E buffer: this thread's lane, the 16 columns that hold its 32 logits
          pb = p_iter & 1;
          mbar_wait(&sb->p_empty[pb], ((p_iter >> 1) & 1) ^ 1);
          tc_fence_after();
          p_taddr = tmem_p + pb * (TCN / 2) + ((uint32_t)(quarter * 32) << 16) + slice * 16;
        }
        if (stamp) p.dbg_ts[1 * 2048 + e_iter * 4 + 1] = clock64();
        // ... while the first is processed
        if (!BWD) {
          if (generic) fwd_chunk<true, PRE, false>(rA, cb, By, c1, c0, dcol, acc, best, best_cb, dval, have_diag);
          else if (PRE && poly) fwd_chunk<false, PRE, true>(rA, cb, By, c1, c0, dcol, acc, best, best_cb, dval, have_diag);
          else fwd_chunk<false, PRE, false>(rA, cb, By, c1, c0, dcol, acc, best, best_cb, dval, have_diag);
        } else {
          if (PRE && poly) bwd_chunk<PRE, SUMS, true>(rA, c1, c0, p_taddr, acc);
          else bwd_chunk<PRE, SUMS, false>(rA, c1, c0, p_taddr, acc);
        }
        if (stamp) p.dbg_ts[1 * 2048 + e_iter * 4 + 2] = clock64();
        tmem_ld_wait_dep(rB);
        // the whole slice is in registers: S buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->s_empty[s_cur]);
        if (t + 1 < t1) {  // next tile's first half in flight while this tile's second half is processed
          mbar_wait(&sb->s_full[s_nxt], s_nxt_ph);
          tc_fence_after();
          tmem_ld16(t_base + s_nxt * TCN, rA);
        }
        if (!BWD) {
          if (generic) fwd_chunk<true, PRE, false>(rB, cb + TC_CW, By, c1, c0, dcol, acc, best, best_cb, dval, have_diag);
          else if (PRE && poly) fwd_chunk<false, PRE, true>(rB, cb + TC_CW, By, c1, c0, dcol, acc, best, best_cb, dval, have_diag);
          else fwd_chunk<false, PRE, false>(rB, cb + TC_CW, By, c1, c0, dcol, acc, best, best_cb, dval, have_diag);
        } else {
          if (PRE && poly) bwd_chunk<PRE, SUMS, true>(rB, c1, c0, p_taddr + TC_CW / 2, acc);
          else bwd_chunk<PRE, SUMS, false>(rB, c1, c0, p_taddr + TC_CW / 2, acc);
          if (SUMS && t == t_part) {
            // columns past By are zero-filled operands: their logit is exactly 0, E exactly 1 (PRE: exp2(0)); they add nothing to
            // the MMA (zero rows of Y) but one each to the sum — take them out again (exact in fp32)
            const int pad = cb + 32 - By;
            if (pad > 0) acc[0] -= (float)(pad > 32 ? 32 : pad) * (PRE ? 1.f : ex2f(c0));
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sb->p_full[pb]);
          ++p_iter;
        }
        if (stamp) p.dbg_ts[1 * 2048 + e_iter * 4 + 3] = clock64();
        ++e_iter;
        if (BWD && pend) {  // rB is dead here, rA already holds the next tile's first half
          drain_d(pend_out, pend_buf, pend_ph, pend_ok, pend_ixs, rB);
          pend = false;
        }
        s_cur = s_nxt;
        s_cur_ph = s_nxt_ph;
      }
      if (!BWD) {
        if (row_ok) {
          const int64_t o = ((int64_t)ch * 4 + slice) * p.Bx + row;
          p.row_e[o] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
          if (p.want_metrics) {
            p.row_max[o] = best;
            p.row_arg[o] = best_cb;
            if (have_diag) p.diag[row] = dval * p.inv_temp;
          }
        }
      } else {
        float* out = p.dx_part + ((int64_t)ch * p.Bx + row) * (SUMS ? 2 * p.F : p.F);
        if (SUMS && row_ok) p.row_e[((int64_t)ch * 4 + slice) * p.Bx + row] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        if (istamp) p.dbg_ts[1 * 2048 + 1024 + it_iter * 4 + 2] = clock64();
        if (!last_item) {  // deferred: read out after the next item's first tile (no wait for MMA2 then)
          pend = true;
          pend_out = out;
          pend_buf = d_buf;
          pend_ph = d_ph;
          pend_ok = row_ok;
          pend_ixs = ixs;
        } else {
          drain_d(out, d_buf, d_ph, row_ok, ixs, rB);
        }
        if (++d_buf == n_dbuf) {
          d_buf = 0;
          d_ph ^= 1;
        }
        if (istamp) p.dbg_ts[1 * 2048 + 1024 + it_iter * 4 + 3] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  if (p.dbg_ts && threadIdx.x == 0) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg_ts[2 * 2048 + 1024 + blockIdx.x * 4 + 2] = (long long)g;
  }
}

// ---------------------------------------------------------------- small deterministic reductions
// sum_j S_ij = inv_temp * X_i . (sum_j Y_j): partial column sums of Y. 512 threads = (F/8) column groups x row lanes, 16-byte
// loads; block b takes rows b*L + l, +L*YS_BLOCKS, ... (fixed order), the row lanes are folded in order through shared memory
constexpr int YS_BLOCKS = 32;
__global__ void __launch_bounds__(512)
tc_ysum_partial_kernel(const __nv_bfloat16* __restrict__ Y, int64_t By, int F, float* __restrict__ part /* [YS_BLOCKS][F] */) {
  __shared__ float sh[64][129];  // [row lane][column], odd pitch
  const int groups = F >> 3, lanes = 512 / groups;  // F = 64: 8 x 64; F = 128: 16 x 32
  const int g = threadIdx.x % groups, l = threadIdx.x / groups;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t r = (int64_t)blockIdx.x * lanes + l; r < By; r += (int64_t)YS_BLOCKS * lanes) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(Y + r * F) + g);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = __bfloat1622float2(h[u]);
      a[2 * u] += f.x;
      a[2 * u + 1] += f.y;
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) sh[l][g * 8 + u] = a[u];
  __syncthreads();
  if (threadIdx.x < F) {
    float t = sh[0][threadIdx.x];
    for (int k = 1; k < lanes; ++k) t += sh[k][threadIdx.x];
    part[blockIdx.x * F + threadIdx.x] = t;
  }
}

// 32 rows x 8 lanes per block. Lane q folds parts q, q+8, ... in order, then the 8 lane results are folded in order (fixed
// association -> deterministic); the arg-max fold is order-independent (max value, then smallest column). The main kernel
// tracks the arg-max per 16-column chunk only: the column inside the winning chunk is resolved here from 16 recomputed dot
// products (first maximum). sum_j S_ij comes from the column sums of Y.
constexpr int RR_LANES = 8;
__global__ void __launch_bounds__(32 * RR_LANES)
tc_rows_reduce_kernel(const float* __restrict__ row_e, const float* __restrict__ row_max, const int32_t* __restrict__ row_arg,
                      int64_t Bx, int n_ct, long long T, int G, const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Y, int64_t By,
                      int F, float inv_temp, const float* __restrict__ ysum_part, float* __restrict__ out_e,
                      float* __restrict__ out_s, int32_t* __restrict__ out_arg, float* __restrict__ out_inv) {
  __shared__ float sh_e[RR_LANES][32], sh_m[RR_LANES][32], sh_d[2 * RR_LANES][32], sh_y[128];
  __shared__ int32_t sh_a[RR_LANES][32];
  const int r = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 32 + r;
  const bool ok = m < Bx;
  if (out_s) {
    for (int c = threadIdx.x; c < F; c += 32 * RR_LANES) {
      float t[YS_BLOCKS];  // all partials in flight, then the fold in block order: one L2 round trip instead of YS_BLOCKS
#pragma unroll
      for (int b = 0; b < YS_BLOCKS; ++b) t[b] = ysum_part[b * F + c];
      float a = 0.f;
#pragma unroll
      for (int b = 0; b < YS_BLOCKS; ++b) a += t[b];
      sh_y[c] = a;
    }
  }
  float e = 0.f, mx = -FLT_MAX;
  int32_t arg = 0x7fffffff;
  if (ok) {
    const int n_parts = 4 * tc_parts_of_row_block((int)(m / TCM), n_ct, T, G);  // 4 column slices per piece of the row block
#pragma unroll 4
    for (int c = q; c < n_parts; c += RR_LANES) {
      const int64_t o = (int64_t)c * Bx + m;
      e += row_e[o];
      if (out_arg) {
        const float tm = row_max[o];
        const int32_t ta = row_arg[o];
        if (tm > mx || (tm == mx && ta < arg)) {
          mx = tm;
          arg = ta;
        }
      }
    }
  }
  sh_e[q][r] = e;
  sh_m[q][r] = mx;
  sh_a[q][r] = arg;
  __syncthreads();
  if (q == 0) {
#pragma unroll
    for (int k = 1; k < RR_LANES; ++k) {
      e += sh_e[k][r];
      const float tm = sh_m[k][r];
      const int32_t ta = sh_a[k][r];
      if (tm > mx || (tm == mx && ta < arg)) {
        mx = tm;
        arg = ta;
      }
    }
    if (ok && out_e) out_e[m] = e;
    if (ok && out_inv) out_inv[m] = 1.f / e;  // the row scale of the backward pass (same expression as tt_simce_finalize)
    sh_a[0][r] = arg;
  }
  __syncthreads();
  // lane q: the part [q*F/8, (q+1)*F/8) of X_m . ysum, and the logits of columns cb + 2q, cb + 2q + 1
  const int cb = sh_a[0][r];
  float ds = 0.f, d0 = -FLT_MAX, d1 = -FLT_MAX;
  if (ok) {
    const __nv_bfloat16* xr = X + m * F;
    if (out_s) {
      const int w = F / RR_LANES;
      for (int k = q * w; k < (q + 1) * w; ++k) ds = fmaf(__bfloat162float(xr[k]), sh_y[k], ds);
    }
    if (out_arg && cb != 0x7fffffff) {
      const int64_t j0 = (int64_t)cb + 2 * q, j1 = j0 + 1;
      const __nv_bfloat16* y0 = Y + j0 * F;
      const __nv_bfloat16* y1 = Y + j1 * F;
      float a0 = 0.f, a1 = 0.f;
      const bool v0 = j0 < By, v1 = j1 < By;
#pragma unroll 4
      for (int k = 0; k < F; k += 8) {  // F is 64 or 128: four iterations' loads (twelve 16-byte vectors) in flight per round
        const uint4 xv = *reinterpret_cast<const uint4*>(xr + k);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        const uint4 yv0 = v0 ? __ldg(reinterpret_cast<const uint4*>(y0 + k)) : z;
        const uint4 yv1 = v1 ? __ldg(reinterpret_cast<const uint4*>(y1 + k)) : z;
        const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xv);
        const __nv_bfloat162* p0 = reinterpret_cast<const __nv_bfloat162*>(&yv0);
        const __nv_bfloat162* p1 = reinterpret_cast<const __nv_bfloat162*>(&yv1);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 xf = __bfloat1622float2(xp[u]), f0 = __bfloat1622float2(p0[u]), f1 = __bfloat1622float2(p1[u]);
          a0 = fmaf(xf.x, f0.x, a0);
          a0 = fmaf(xf.y, f0.y, a0);
          a1 = fmaf(xf.x, f1.x, a1);
          a1 = fmaf(xf.y, f1.y, a1);
        }
      }
      if (v0) d0 = a0;
      if (v1) d1 = a1;
    }
  }
  sh_e[q][r] = ds;
  sh_d[2 * q][r] = d0;
  sh_d[2 * q + 1][r] = d1;
  __syncthreads();
  if (q == 0 && ok) {
    if (out_s) {
      float t = sh_e[0][r];
#pragma unroll
      for (int k = 1; k < RR_LANES; ++k) t += sh_e[k][r];
      out_s[m] = t * inv_temp;
    }
    if (out_arg) {
      int jj = 0;
      float bm = sh_d[0][r];
#pragma unroll
      for (int k = 1; k < 2 * RR_LANES; ++k) {
        const float t = sh_d[k][r];
        if (t > bm) {
          bm = t;
          jj = k;
        }
      }
      out_arg[m] = cb == 0x7fffffff ? 0 : cb + jj;
    }
  }
}

// Ys[j,:] = bf16(gscale * inv_y[j] * Y[j,:]): the column scales of dS folded into the second operand of the backward MMA
__global__ void __launch_bounds__(256)
tc_scale_rows_kernel(const __nv_bfloat16* __restrict__ Y, const float* __restrict__ inv_y, float gscale, int64_t By, int F,
                     __nv_bfloat16* __restrict__ Ys) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= By * F) return;
  const float sc = gscale * inv_y[i / F];
  const uint4 v = *reinterpret_cast<const uint4*>(Y + i);
  const __nv_bfloat162* in = reinterpret_cast<const __nv_bfloat162*>(&v);
  uint4 o;
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 f = __bfloat1622float2(in[u]);
    out[u] = __floats2bfloat162_rn(f.x * sc, f.y * sc);
  }
  *reinterpret_cast<uint4*>(Ys + i) = o;
}

__global__ void tc_ysum_final_kernel(const float* __restrict__ part, int F, float* __restrict__ ysum) {
  const int c = threadIdx.x;
  if (c >= F) return;
  float a = 0.f;
  for (int b = 0; b < YS_BLOCKS; ++b) a += part[b * F + c];
  ysum[c] = a;
}

// dX = sum of the chunk partials (fixed order) - t_off * sum_j Y_j - (t_diag - t_off) * Y[diag_offset + row]
__global__ void __launch_bounds__(256)
tc_dx_reduce_kernel(const float* __restrict__ part, int64_t n, int n_ct, long long T, int G, int F, const __nv_bfloat16* __restrict__ Y,
                    int64_t diag_offset, float t_off2, float t_dd2, const float* __restrict__ ysum, float* __restrict__ dX) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const int64_t row = i / F;
  const int n_parts = tc_parts_of_row_block((int)(row / TCM), n_ct, T, G);
  float4 a = *reinterpret_cast<const float4*>(part + i);
  if (n_parts <= 4) {  // all loads first, then the adds in part order (see tc_finish_kernel)
    float4 q[3];
#pragma unroll
    for (int c = 1; c < 4; ++c) q[c - 1] = *reinterpret_cast<const float4*>(part + (int64_t)(c < n_parts ? c : 0) * n + i);
#pragma unroll
    for (int c = 1; c < 4; ++c)
      if (c < n_parts) f4_add(a, q[c - 1]);
  } else {
    for (int c = 1; c < n_parts; ++c) f4_add(a, *reinterpret_cast<const float4*>(part + (int64_t)c * n + i));
  }
  const int col = (int)(i - row * F);
  const uint2 yv = *reinterpret_cast<const uint2*>(Y + (diag_offset + row) * F + col);
  const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.x));
  const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.y));
  a.x = fmaf(-t_dd2, y01.x, a.x);
  a.y = fmaf(-t_dd2, y01.y, a.y);
  a.z = fmaf(-t_dd2, y23.x, a.z);
  a.w = fmaf(-t_dd2, y23.y, a.w);
  if (ysum) {
    const float4 ys = *reinterpret_cast<const float4*>(ysum + col);
    a.x = fmaf(-t_off2, ys.x, a.x);
    a.y = fmaf(-t_off2, ys.y, a.y);
    a.z = fmaf(-t_off2, ys.z, a.z);
    a.w = fmaf(-t_off2, ys.w, a.w);
  }
  *reinterpret_cast<float4*>(dX + i) = a;
}

// ---------------------------------------------------------------- host side
static bool make_map(CUtensorMap* map, const void* ptr, int64_t rows, int F) {
  return make_map_bf16(map, ptr, rows, F, F, 64, TCM);
}

struct TcPlan {
  int n_row_blocks, n_col_tiles, grid, max_parts;
  long long total_tiles;
};

// One contiguous tile range per CTA (see tc_range_begin): a CTA works through at most a few long items, a row block is split over
// at most max_parts CTAs (so that many partial results per row), and every CTA gets floor or ceil of T/G tiles.
static TcPlan make_plan(int64_t Bx, int64_t By) {
  TcPlan pl;
  pl.n_row_blocks = (int)((Bx + TCM - 1) / TCM);
  pl.n_col_tiles = (int)((By + TCN - 1) / TCN);
  pl.total_tiles = (long long)pl.n_row_blocks * pl.n_col_tiles;
  const int sms = sm_count();
  pl.grid = pl.total_tiles < sms ? (int)pl.total_tiles : sms;
  int mp = 1;
  for (int rb = 0; rb < pl.n_row_blocks; ++rb) {
    const int n = tc_parts_of_row_block(rb, pl.n_col_tiles, pl.total_tiles, pl.grid);
    if (n > mp) mp = n;
  }
  pl.max_parts = mp;
  return pl;
}

static void tc_buffers(int F, bool bwd, int& n_stages, int& n_xbuf) {
  // backward stages hold Y and Ys (2 tiles) and 64 KB go to the two E buffers: stay under 227 KB
  const bool tight = bwd && F > 64;
  n_stages = !bwd ? 4 : (tight ? 3 : 6);  // backward stages hold Y and Ys: F = 64: 32 + 6 * 32 KB, F = 128: 32 + 3 * 64 KB
  n_xbuf = tight ? 1 : 2;
}
static size_t tc_smem_bytes(int F, bool bwd) {
  const int tile = (F >> 6) * ATOM_BYTES;
  int ns, nx;
  tc_buffers(F, bwd, ns, nx);
  return 1024 + (size_t)tile * (nx + ns * (bwd ? 2 : 1)) + sizeof(TcSmem) + 64;
}

int64_t simce_tc_workspace_bytes(int64_t Bx, int64_t By, int F) {
  const int64_t B = Bx > By ? Bx : By;
  TcPlan a = make_plan(Bx, By), b = make_plan(By, Bx);
  const int64_t chunks = a.max_parts > b.max_parts ? a.max_parts : b.max_parts;
  const int64_t fwd = 3 * ((4 * chunks * B * 4 + 255) / 256 * 256) + YS_BLOCKS * 128 * 4 + 256;
  const int64_t bwd = chunks * B * F * 4 + 256 + (B * F * 2 + 256) + YS_BLOCKS * 128 * 4 + 128 * 4 + 256;
  return (fwd > bwd ? fwd : bwd) + 1024;
}

static int g_tc_force_manual = 0;
static int g_tc_poly_fwd = 0, g_tc_poly_bwd = 1;  // measured at B = 8192: P1 26.5 us without / 27.1 with, P2 39.6 -> 35.7, P3 32.7 -> 29.9
static long long* g_tc_dbg_ts = nullptr;

// rows statistics of S = X Y^T (bf16 inputs).  want_metrics: also max/argmax/diag.
// sums_only: only the row sums (and their reciprocals, out_inv) are reduced here; simce_tc_rows_metrics finishes the job later
// from the partials this call leaves in `workspace`.
int simce_tc_rows(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                  float* row_sumexp, float* row_sum_s, float* diag, int32_t* row_argmax, int want_metrics, void* workspace,
                  cudaStream_t s, int sums_only, float* out_inv, int pre) {
  TcPlan pl = make_plan(Bx, By);
  TcParams p = {};
  p.X = (const __nv_bfloat16*)X;
  p.Y = (const __nv_bfloat16*)Y;
  p.Bx = Bx; p.By = By; p.F = F;
  p.n_col_tiles = pl.n_col_tiles; p.total_tiles = pl.total_tiles;
  const float log2e = 1.4426950408889634f;
  p.c1 = inv_temp * log2e; p.c0 = -inv_temp * log2e; p.inv_temp = pre ? 0.6931471805599453f : inv_temp;
  p.diag_offset = diag_offset;
  const int64_t stride = (4 * pl.max_parts * Bx * 4 + 255) / 256 * 256;
  char* w = (char*)workspace;
  p.row_e = (float*)w; p.row_max = (float*)(w + stride); p.row_arg = (int32_t*)(w + 2 * stride);
  float* ysum_part = (float*)(w + 3 * stride);
  p.diag = diag; p.want_metrics = want_metrics; p.dbg_ts = want_metrics ? g_tc_dbg_ts : nullptr;
  tc_buffers(F, false, p.n_stages, p.n_xbuf);
  p.n_dbuf = 1;
  p.n_sbuf = TC_SBUF;
  p.poly = pre ? g_tc_poly_fwd : 0;
  CUtensorMap mx, my;
  p.use_tma = (!g_tc_force_manual && make_map(&mx, X, Bx, F) && make_map(&my, Y, By, F)) ? 1 : 0;
  if (!p.use_tma) { memset(&mx, 0, sizeof(mx)); memset(&my, 0, sizeof(my)); }
  const size_t smem = tc_smem_bytes(F, false);
  if (pre) {
    TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<TC_FWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TT_CUDA(launch_pdl(simce_tc_kernel<TC_FWD, true>, dim3(pl.grid), dim3(TC_THREADS), smem, s, mx, my, my, p));
    TT_LAUNCH_CHECK("simce_tc_kernel<fwd,pre>");
    return TT_OK;
  }
  TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<TC_FWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simce_tc_kernel<TC_FWD, false><<<pl.grid, TC_THREADS, smem, s>>>(mx, my, my, p);
  if (sums_only) {
    tc_rows_reduce_kernel<<<ceil_div(Bx, 32), 32 * RR_LANES, 0, s>>>(p.row_e, p.row_max, p.row_arg, Bx, pl.n_col_tiles, pl.total_tiles,
                                                                     pl.grid, p.X, p.Y, By, F, inv_temp, nullptr, row_sumexp, nullptr,
                                                                     nullptr, out_inv);
    TT_LAUNCH_CHECK("simce_tc_rows(sums)");
    return TT_OK;
  }
  if (row_sum_s) tc_ysum_partial_kernel<<<YS_BLOCKS, 512, 0, s>>>(p.Y, By, F, ysum_part);
  tc_rows_reduce_kernel<<<ceil_div(Bx, 32), 32 * RR_LANES, 0, s>>>(p.row_e, p.row_max, p.row_arg, Bx, pl.n_col_tiles, pl.total_tiles, pl.grid, p.X, p.Y, By, F,
                                                                   inv_temp, ysum_part, row_sumexp, row_sum_s,
                                                                   want_metrics ? row_argmax : nullptr, out_inv);
  TT_LAUNCH_CHECK("simce_tc_rows");
  return TT_OK;
}

// The part of the forward statistics that only the loss VALUE and the metrics need (sum_j S_ij, arg-max), from the partials a
// sums_only simce_tc_rows call left in `workspace`: off the critical path of the train step.
int simce_tc_rows_metrics(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, float* row_sum_s,
                          int32_t* row_argmax, void* workspace, cudaStream_t s) {
  TcPlan pl = make_plan(Bx, By);
  const int64_t stride = (4 * pl.max_parts * Bx * 4 + 255) / 256 * 256;
  char* w = (char*)workspace;
  const float* row_e = (const float*)w;
  const float* row_max = (const float*)(w + stride);
  const int32_t* row_arg = (const int32_t*)(w + 2 * stride);
  float* ysum_part = (float*)(w + 3 * stride);
  if (row_sum_s) tc_ysum_partial_kernel<<<YS_BLOCKS, 512, 0, s>>>((const __nv_bfloat16*)Y, By, F, ysum_part);
  tc_rows_reduce_kernel<<<ceil_div(Bx, 32), 32 * RR_LANES, 0, s>>>(row_e, row_max, row_arg, Bx, pl.n_col_tiles, pl.total_tiles, pl.grid,
                                                                   (const __nv_bfloat16*)X, (const __nv_bfloat16*)Y, By, F, inv_temp,
                                                                   ysum_part, nullptr, row_sum_s, row_argmax, nullptr);
  TT_LAUNCH_CHECK("simce_tc_rows_metrics");
  return TT_OK;
}

int simce_tc_bwd(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                 const float* inv_x, const float* inv_y, float gscale, float tgt_diag, float tgt_off, float* dX, void* workspace,
                 cudaStream_t s) {
  TcPlan pl = make_plan(Bx, By);
  TcParams p = {};
  p.X = (const __nv_bfloat16*)X;
  p.Y = (const __nv_bfloat16*)Y;
  p.Bx = Bx; p.By = By; p.F = F;
  p.n_col_tiles = pl.n_col_tiles; p.total_tiles = pl.total_tiles;
  const float log2e = 1.4426950408889634f;
  p.c1 = inv_temp * log2e; p.c0 = -inv_temp * log2e; p.inv_temp = inv_temp;
  p.diag_offset = diag_offset;
  p.inv_x = inv_x; p.gscale = gscale;
  const float t_diag2 = 2.f * gscale * tgt_diag, t_off2 = 2.f * gscale * tgt_off;
  // workspace: [chunk partials of dX][Ys][column-sum partials][column sums]
  char* w = (char*)workspace;
  p.dx_part = (float*)w;
  w += ((int64_t)pl.max_parts * Bx * F * 4 + 255) / 256 * 256;
  __nv_bfloat16* Ys = (__nv_bfloat16*)w;
  w += (By * F * 2 + 255) / 256 * 256;
  float* ysum_part = (float*)w;
  float* ysum = ysum_part + YS_BLOCKS * 128;
  p.Ys = Ys;
  p.dbg_ts = g_tc_dbg_ts;
  tc_buffers(F, true, p.n_stages, p.n_xbuf);
  p.n_dbuf = 1;
  p.n_sbuf = F == 64 ? 2 : 1;  // S buffers + 2F accumulator columns + 128 columns of E must fit 512
  tc_scale_rows_kernel<<<ceil_div(By * F / 8, 256), 256, 0, s>>>(p.Y, inv_y, gscale, By, F, Ys);
  const bool smooth = tgt_off != 0.f;
  if (smooth) {
    tc_ysum_partial_kernel<<<YS_BLOCKS, 512, 0, s>>>(p.Y, By, F, ysum_part);
    tc_ysum_final_kernel<<<1, 128, 0, s>>>(ysum_part, F, ysum);
  }
  CUtensorMap mx, my, mys;
  p.use_tma = (!g_tc_force_manual && make_map(&mx, X, Bx, F) && make_map(&my, Y, By, F) && make_map(&mys, Ys, By, F)) ? 1 : 0;
  if (!p.use_tma) { memset(&mx, 0, sizeof(mx)); memset(&my, 0, sizeof(my)); memset(&mys, 0, sizeof(mys)); }
  const size_t smem = tc_smem_bytes(F, true);
  TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<TC_BWD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simce_tc_kernel<TC_BWD, false><<<pl.grid, TC_THREADS, smem, s>>>(mx, my, mys, p);
  tc_dx_reduce_kernel<<<ceil_div(Bx * F / 4, 256), 256, 0, s>>>(p.dx_part, Bx * F, pl.n_col_tiles, pl.total_tiles, pl.grid, F, p.Y, diag_offset, t_off2,
                                                                  t_diag2 - t_off2, smooth ? ysum : nullptr, dX);
  TT_LAUNCH_CHECK("simce_tc_bwd");
  return TT_OK;
}


// ---------------------------------------------------------------- three-pass train step (pre-scaled X operand)
// The fused step runs THREE logits passes instead of four:
//   P1  rows of S        X = c1*N,  Y = C            -> sum_j E_ij                        (simce_tc_kernel<TC_FWD, pre>)
//   P2  rows of S^T      X = c1*C,  Y = [N | Ns]     -> sum_i E_ij AND the raw halves of dC (simce_tc_kernel<TC_BWD_SUMS, pre>)
//   P3  rows of S again  X = c1*N,  Y = [C | Cs]     -> dN                                (simce_tc_kernel<TC_BWD, pre>)
// with Ns = diag(g / rowsum) N (needs P1), Cs = diag(g / colsum) C (needs P2): the transposed gradient pass is the column-
// statistics pass, its own row scale being applied by the reduce kernel once the pass has finished. E = exp2(c1 * s) here
// (unshifted); the public sums stay "shifted" (sum exp(s/T - 1/T)) like everywhere else: k = exp(-1/T).
//
// tc_finish_kernel: per row, in fixed order: rs = sum of the E partials; row_sumexp = k rs, inv = 1 / row_sumexp;
//   dX = (g / rs) sum_parts(E.Y) + sum_parts(E.Ys) - t_off sum_j Y_j - (t_diag - t_off) Y_dcol      (if dx_part)
//   Xs = bf16((g / rs) X_unit)   — the column-scaled operand of the NEXT pass                        (if Xs_out)
__global__ void __launch_bounds__(256)
tc_finish_kernel(const float* __restrict__ row_e, const float* __restrict__ dx_part, int64_t Bx, int n_ct, long long T, int G, int F,
                 const __nv_bfloat16* __restrict__ Y, int64_t diag_offset, float t_off2, float t_dd2, const float* __restrict__ ysum,
                 float gscale, float k_shift, float* __restrict__ dX, float* __restrict__ sumexp_out, float* __restrict__ inv_out,
                 const __nv_bfloat16* __restrict__ X_unit, __nv_bfloat16* __restrict__ Xs_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= Bx * F) return;
  const int64_t row = i / F;
  const int col = (int)(i - row * F);
  const int n_parts = tc_parts_of_row_block((int)(row / TCM), n_ct, T, G);
  // The kernel is one L2 round trip long if every load leaves before the first add: up to four parts (the usual split of a row
  // block) are fetched into registers in one go and then added in part order — the same order as the plain loops below.
  float rs = 0.f;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (n_parts <= 4) {
    float e[16];
    float4 pa[4], pb[4];
#pragma unroll
    for (int c = 0; c < 16; ++c) e[c] = c < 4 * n_parts ? row_e[(int64_t)c * Bx + row] : 0.f;
    if (dx_part) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* base = dx_part + ((int64_t)(c < n_parts ? c : 0) * Bx + row) * (2 * F) + col;
        pa[c] = *reinterpret_cast<const float4*>(base);
        pb[c] = *reinterpret_cast<const float4*>(base + F);
      }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < 4 * n_parts) rs += e[c];
    if (dx_part) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < n_parts) {
          f4_add(a, pa[c]);
          f4_add(b, pb[c]);
        }
    }
  } else {
    for (int c = 0; c < 4 * n_parts; ++c) rs += row_e[(int64_t)c * Bx + row];
    if (dx_part) {
      for (int c = 0; c < n_parts; ++c) {
        const float* base = dx_part + ((int64_t)c * Bx + row) * (2 * F) + col;
        f4_add(a, *reinterpret_cast<const float4*>(base));
        f4_add(b, *reinterpret_cast<const float4*>(base + F));
      }
    }
  }
  const float ixs = gscale / rs;
  if (col == 0) {
    const float se = k_shift * rs;
    if (sumexp_out) sumexp_out[row] = se;
    if (inv_out) inv_out[row] = 1.f / se;
  }
  if (dx_part) {
    a.x = fmaf(ixs, a.x, b.x);
    a.y = fmaf(ixs, a.y, b.y);
    a.z = fmaf(ixs, a.z, b.z);
    a.w = fmaf(ixs, a.w, b.w);
    const uint2 yv = *reinterpret_cast<const uint2*>(Y + (diag_offset + row) * F + col);
    const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.x));
    const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.y));
    a.x = fmaf(-t_dd2, y01.x, a.x);
    a.y = fmaf(-t_dd2, y01.y, a.y);
    a.z = fmaf(-t_dd2, y23.x, a.z);
    a.w = fmaf(-t_dd2, y23.y, a.w);
    if (ysum) {
      const float4 ys = *reinterpret_cast<const float4*>(ysum + col);
      a.x = fmaf(-t_off2, ys.x, a.x);
      a.y = fmaf(-t_off2, ys.y, a.y);
      a.z = fmaf(-t_off2, ys.z, a.z);
      a.w = fmaf(-t_off2, ys.w, a.w);
    }
    *reinterpret_cast<float4*>(dX + i) = a;
  }
  if (Xs_out) {
    const uint2 xv = *reinterpret_cast<const uint2*>(X_unit + i);
    const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv.x));
    const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xv.y));
    *reinterpret_cast<uint2*>(Xs_out + i) = pack_bf16x4(make_float4(x01.x * ixs, x01.y * ixs, x23.x * ixs, x23.y * ixs));
  }
}

// Ys[j,:] = bf16(g * k / sumexp[j] * Y[j,:]) from the (exchanged, shifted) sums: the multi-GPU form of the Xs output above
__global__ void __launch_bounds__(256)
tc_scale_rows_sum_kernel(const __nv_bfloat16* __restrict__ Y, const float* __restrict__ sumexp, float gk, int64_t By, int F,
                         __nv_bfloat16* __restrict__ Ys) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= By * F) return;
  const float sc = gk / sumexp[i / F];
  const uint4 v = *reinterpret_cast<const uint4*>(Y + i);
  const __nv_bfloat162* in = reinterpret_cast<const __nv_bfloat162*>(&v);
  uint4 o;
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 f = __bfloat1622float2(in[u]);
    out[u] = __floats2bfloat162_rn(f.x * sc, f.y * sc);
  }
  *reinterpret_cast<uint4*>(Ys + i) = o;
}

struct Tc3Ws {  // workspace of one gradient pass
  float* dx_part;
  float* ysum_part;
  float* ysum;
  float* row_e;
};
static int64_t tc3_a256(int64_t x) { return (x + 255) / 256 * 256; }
static Tc3Ws tc3_carve(void* workspace, int64_t Bx, int F, int max_parts) {
  char* w = (char*)workspace;
  Tc3Ws r;
  r.dx_part = (float*)w;
  w += tc3_a256((int64_t)max_parts * Bx * 2 * F * 4);
  r.ysum_part = (float*)w;
  w += tc3_a256((int64_t)YS_BLOCKS * 128 * 4);
  r.ysum = (float*)w;
  w += tc3_a256(128 * 4);
  r.row_e = (float*)w;
  return r;
}
int64_t simce_tc_grad_workspace_bytes(int64_t Bx, int64_t By, int F) {
  TcPlan pl = make_plan(Bx, By);
  return tc3_a256((int64_t)pl.max_parts * Bx * 2 * F * 4) + tc3_a256((int64_t)YS_BLOCKS * 128 * 4) + tc3_a256(128 * 4) +
         tc3_a256((int64_t)4 * pl.max_parts * Bx * 4) + 1024;
}

// P1: statistics pass; partials stay in `workspace` (layout of simce_tc_rows) for simce_tc_rows_finish / simce_tc_rows_metrics
int simce_tc_rows_pre(const void* Xs, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset, float* diag,
                      void* workspace, cudaStream_t s) {
  return simce_tc_rows(Xs, Y, Bx, By, F, inv_temp, diag_offset, nullptr, nullptr, diag, nullptr, 1, workspace, s, 0, nullptr, 1);
}

int simce_tc_rows_finish(const void* workspace, int64_t Bx, int64_t By, int F, float inv_temp, float* row_sumexp, float* inv_row,
                         const void* X_unit, float gscale, void* Xs_out, cudaStream_t s) {
  TcPlan pl = make_plan(Bx, By);
  const float* row_e = (const float*)workspace;
  TT_CUDA(launch_pdl(tc_finish_kernel, dim3(ceil_div(Bx * F / 4, 256)), dim3(256), 0, s, row_e, (const float*)nullptr, Bx, pl.n_col_tiles,
                     pl.total_tiles, pl.grid, F, (const __nv_bfloat16*)nullptr, (int64_t)0, 0.f, 0.f, (const float*)nullptr, gscale,
                     expf(-inv_temp), (float*)nullptr, row_sumexp, inv_row, (const __nv_bfloat16*)X_unit, (__nv_bfloat16*)Xs_out));
  TT_LAUNCH_CHECK("tc_finish_kernel(rows)");
  return TT_OK;
}

// P2 / P3. inv_x == NULL: deferred row scale (TC_BWD_SUMS), else inv_x are the (shifted) reciprocal row sums (TC_BWD).
int simce_tc_grad_pre(const void* Xs, const void* Y, const void* Ys, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                      const float* inv_x, float gscale, int smooth, void* workspace, cudaStream_t s) {
  TcPlan pl = make_plan(Bx, By);
  Tc3Ws w = tc3_carve(workspace, Bx, F, pl.max_parts);
  TcParams p = {};
  p.X = (const __nv_bfloat16*)Xs;
  p.Y = (const __nv_bfloat16*)Y;
  p.Ys = (const __nv_bfloat16*)Ys;
  p.Bx = Bx; p.By = By; p.F = F;
  p.n_col_tiles = pl.n_col_tiles; p.total_tiles = pl.total_tiles;
  p.c1 = 1.f; p.c0 = 0.f; p.inv_temp = 0.6931471805599453f;
  p.diag_offset = diag_offset;
  p.inv_x = inv_x;
  p.gscale = gscale * expf(-inv_temp);  // E is unshifted here, the public reciprocals belong to the shifted sums
  p.dx_part = w.dx_part;
  p.row_e = w.row_e;
  p.poly = g_tc_poly_bwd;
  p.dbg_ts = g_tc_dbg_ts;
  tc_buffers(F, true, p.n_stages, p.n_xbuf);
  p.n_dbuf = 1;
  p.n_sbuf = F == 64 ? 2 : 1;
  if (smooth) {
    tc_ysum_partial_kernel<<<YS_BLOCKS, 512, 0, s>>>(p.Y, By, F, w.ysum_part);
    tc_ysum_final_kernel<<<1, 128, 0, s>>>(w.ysum_part, F, w.ysum);
  }
  CUtensorMap mx, my, mys;
  p.use_tma = (!g_tc_force_manual && make_map(&mx, Xs, Bx, F) && make_map(&my, Y, By, F) && make_map(&mys, Ys, By, F)) ? 1 : 0;
  if (!p.use_tma) { memset(&mx, 0, sizeof(mx)); memset(&my, 0, sizeof(my)); memset(&mys, 0, sizeof(mys)); }
  const size_t smem = tc_smem_bytes(F, true);
  if (inv_x) {
    TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<TC_BWD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TT_CUDA(launch_pdl(simce_tc_kernel<TC_BWD, true>, dim3(pl.grid), dim3(TC_THREADS), smem, s, mx, my, mys, p));
  } else {
    TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<TC_BWD_SUMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TT_CUDA(launch_pdl(simce_tc_kernel<TC_BWD_SUMS, true>, dim3(pl.grid), dim3(TC_THREADS), smem, s, mx, my, mys, p));
  }
  TT_LAUNCH_CHECK("simce_tc_kernel<bwd,pre>");
  return TT_OK;
}

int simce_tc_grad_finish(const void* workspace, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset, const void* Y,
                         int deferred, float gscale, float tgt_diag, float tgt_off, float* dX, float* sumexp_out, float* inv_out,
                         const void* X_unit, void* Xs_out, cudaStream_t s) {
  TcPlan pl = make_plan(Bx, By);
  Tc3Ws w = tc3_carve(const_cast<void*>(workspace), Bx, F, pl.max_parts);
  const float t_diag2 = 2.f * gscale * tgt_diag, t_off2 = 2.f * gscale * tgt_off;
  const bool smooth = tgt_off != 0.f;
  if (deferred) {
    TT_CUDA(launch_pdl(tc_finish_kernel, dim3(ceil_div(Bx * F / 4, 256)), dim3(256), 0, s, (const float*)w.row_e, (const float*)w.dx_part, Bx,
                       pl.n_col_tiles, pl.total_tiles, pl.grid, F, (const __nv_bfloat16*)Y, diag_offset, t_off2, t_diag2 - t_off2,
                       (const float*)(smooth ? w.ysum : nullptr), gscale, expf(-inv_temp), dX, sumexp_out, inv_out,
                       (const __nv_bfloat16*)X_unit, (__nv_bfloat16*)Xs_out));
  } else {
    TT_CUDA(launch_pdl(tc_dx_reduce_kernel, dim3(ceil_div(Bx * F / 4, 256)), dim3(256), 0, s, (const float*)w.dx_part, Bx * F, pl.n_col_tiles,
                       pl.total_tiles, pl.grid, F, (const __nv_bfloat16*)Y, diag_offset, t_off2, t_diag2 - t_off2,
                       (const float*)(smooth ? w.ysum : nullptr), dX));
  }
  TT_LAUNCH_CHECK("simce_tc_grad_finish");
  return TT_OK;
}

int simce_tc_scale_rows(const void* Y_unit, const float* sumexp, float gscale, float inv_temp, int64_t By, int F, void* Ys_out,
                        cudaStream_t s) {
  tc_scale_rows_sum_kernel<<<ceil_div(By * F / 8, 256), 256, 0, s>>>((const __nv_bfloat16*)Y_unit, sumexp, gscale * expf(-inv_temp), By, F,
                                                                      (__nv_bfloat16*)Ys_out);
  TT_LAUNCH_CHECK("tc_scale_rows_sum_kernel");
  return TT_OK;
}

void simce_tc_set_force_manual(int v) { g_tc_force_manual = v; }
void simce_tc_set_debug(int v) {  // A/B knob of the pre-scaled passes: bit 0 = polynomial share in P1 (default off), bit 1 = NO
  g_tc_poly_fwd = (v & 1) ? 1 : 0;  // polynomial share in P2 / P3 (default on)
  g_tc_poly_bwd = (v & 2) ? 0 : 1;
}
void simce_tc_set_timeline(long long* buf) { g_tc_dbg_ts = buf; }

}  // namespace tt
