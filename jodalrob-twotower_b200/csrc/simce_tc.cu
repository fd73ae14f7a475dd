// Similarity logits + symmetric softmax-CE on the 5th-gen tensor cores (tcgen05) — bf16 path.
//
// One CTA owns a 128-row block of X and walks a chunk of 128-column tiles of Y:
//     S_tile[128x128] = X_blk[128xF] · Y_tile[128xF]^T      tcgen05.mma, fp32 accumulator in TMEM
// The logits never leave the SM: the epilogue warps pull the tile out of TMEM with tcgen05.ld (one row per
// thread, so row reductions need no shuffles), take exp2 of the shifted logits and
//   forward : keep per-row running sum(exp), sum(S), max/argmax and the diagonal;
//   backward: build dS = E*(inv_x_i + inv_y_j) - 2*target in registers, park it as bf16 in shared memory
//             (K-major, 128B-swizzled) and feed it straight back to the tensor core:
//             dX_blk[128xF] += dS_tile[128x128] · Y_tile[128xF]   accumulated in TMEM across the chunk.
// Column statistics / the other tower's gradient come from running the same kernel on the transposed problem
// (X and Y swapped), so no cross-CTA column reduction and no atomics: everything is deterministic.
//
// Warp roles (608 threads): warps 0-15 epilogue (TMEM lane quarter == warp % 4, 32-column slice == warp / 4), warp 16 MMA
// issuer + TMEM allocator, warps 9-10 operand loaders (TMA: one elected lane issues cp.async.bulk.tensor with 128B swizzle;
// the generic-proxy fallback loads 16-byte chunks and applies the same swizzle by hand).
// Pipelines: Y tiles through a 3-stage smem ring (full/empty mbarriers), S through 2 TMEM buffers, dS through
// 2 smem buffers; tcgen05.commit arrives on the barriers when the MMAs that read a buffer have retired.
#include "tc_common.cuh"
#include <float.h>

namespace tt {

constexpr int TCM = 128;          // rows per CTA block (UMMA M)
constexpr int TCN = 128;          // columns per tile  (UMMA N)
constexpr int TC_STAGES = 4;      // Y smem ring
constexpr int TC_SBUF = 3;        // S accumulators in TMEM (3 x 128 columns) so the MMA runs two tiles ahead of the epilogue
constexpr int TC_EPI_WARPS = 16;  // four per SM sub-partition (the epilogue is MUFU/latency bound): rows = TMEM lane quarter
                                  // (warp % 4), columns = the 32-column slice warp / 4 of every 128-column tile
constexpr int TC_MMA_WARP = 16;
constexpr int TC_THREADS = 608;   // 16 epilogue + 1 mma + 2 loader warps
constexpr int TC_LOADER_THREADS = 64;

struct TcParams {
  const __nv_bfloat16* X;
  const __nv_bfloat16* Y;
  int64_t Bx, By;
  int F;                 // 64 or 128
  int n_col_tiles;       // ceil(By / 128)
  int tiles_per_chunk;
  int n_chunks;
  int n_items;           // row blocks * chunks
  float c1, c0;          // e = exp2(s_raw * c1 + c0):  c1 = inv_temp*log2e, c0 = -inv_temp*log2e
  float inv_temp;
  int64_t diag_offset;
  int use_tma;
  int n_stages;          // Y smem ring depth (<= TC_STAGES)
  int n_xbuf;            // X row-block buffers (1 or 2)
  long long* dbg_ts;     // optional timeline buffer (block 0): [role][event] clock64 stamps, see scripts/time_simce.py
  int debug;             // profiling knobs (tt_debug_simce_knobs): 2 = no TMEM loads, 8 = no TMA after the ring fill, 16 = no MMAs
  // forward outputs: partials per (chunk, row)
  float* row_e;
  float* row_s;
  float* row_max;
  int32_t* row_arg;
  float* diag;
  int want_metrics;
  int want_sum_s;        // also accumulate sum_j S_ij (metrics / label smoothing)
  // backward
  const float* inv_x;    // [Bx] (unscaled)
  const float* inv_y;    // [By]
  float gscale, tgt_diag2, tgt_off2;  // 2*gscale*target values
  float* dx_part;        // [n_chunks][Bx][F]
};

struct TcSmem {  // carved from dynamic smem after the 1024-aligned tiles
  uint64_t x_full[2], x_empty[2];  // the X row block is double-buffered so the next work item's block loads early
  uint64_t y_full[TC_STAGES], y_empty[TC_STAGES];
  uint64_t s_full[TC_SBUF], s_empty[TC_SBUF];
  uint64_t p_full[2], p_empty[2];
  uint64_t d_full, d_empty;
  uint32_t tmem_base;
  alignas(16) float iy[TC_SBUF * TCN];
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
// row statistics of 32 logits (forward).  GENERIC handles ragged column tiles and the diagonal.
template <bool GENERIC, bool SUMS>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&r)[32], int64_t cbase, const TcParams& p, int64_t dcol, float (&acc)[4],
                                          float (&accs)[4], float& best, int32_t& best_j, float& dval, bool& have_diag) {
  float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float raw = __uint_as_float(r[j]);
    if (GENERIC) {
      const int64_t col = cbase + j;
      if (col < p.By) {
        acc[j & 3] += ex2f(fmaf(raw, p.c1, p.c0));
        if (SUMS) accs[j & 3] += raw;
        m4[j & 3] = fmaxf(m4[j & 3], raw);
        if (col == dcol) {
          dval = raw;
          have_diag = true;
        }
      }
    } else {
      acc[j & 3] += ex2f(fmaf(raw, p.c1, p.c0));
      if (SUMS) accs[j & 3] += raw;
      m4[j & 3] = fmaxf(m4[j & 3], raw);
    }
  }
  const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  if (m > best) {  // rare after the first few chunks: locate the first column that attains the chunk maximum
    best = m;
    int32_t jj = 31;
#pragma unroll
    for (int j = 31; j >= 0; --j)
      if (__uint_as_float(r[j]) == m && (!GENERIC || cbase + j < p.By)) jj = j;
    best_j = (int32_t)(cbase + jj);
  }
}

// dS for 32 logits -> bf16, written as four 16-byte chunks of this row into the swizzled K-major P tile
template <bool GENERIC>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&r)[32], int64_t cbase, int c_in_tile, const TcParams& p, int64_t dcol,
                                          bool row_ok, float ixs, const float* __restrict__ s_iy, unsigned char* prow, int rsw) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // 8 columns -> one 16-byte chunk
    const float4 a = *reinterpret_cast<const float4*>(s_iy + c_in_tile + q * 8);
    const float4 b = *reinterpret_cast<const float4*>(s_iy + c_in_tile + q * 8 + 4);
    const float iy[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float d[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = q * 8 + u;
      const float e = ex2f(fmaf(__uint_as_float(r[j]), p.c1, p.c0));
      float v = fmaf(e, ixs + iy[u], -p.tgt_off2);
      if (GENERIC) {
        if (cbase + j == dcol) v = fmaf(e, ixs + iy[u], -p.tgt_diag2);
        if (cbase + j >= p.By || !row_ok) v = 0.f;
      }
      d[u] = v;
    }
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      __nv_bfloat162 t = __floats2bfloat162_rn(d[2 * u], d[2 * u + 1]);
      w[u] = *reinterpret_cast<uint32_t*>(&t);
    }
    const int cc = ((c_in_tile & 63) >> 3) + q;  // 16-byte chunk index inside the 64-column atom
    *reinterpret_cast<uint4*>(prow + ((cc ^ rsw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
simce_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NA = p.F >> 6;                      // swizzle atoms along K (1 or 2)
  const int tile_bytes = NA * ATOM_BYTES;       // X / Y tile
  const int NST = p.n_stages;
  unsigned char* sX = base;                     // n_xbuf tiles (double-buffered across work items when they fit)
  unsigned char* sY = sX + p.n_xbuf * tile_bytes;  // NST tiles
  unsigned char* sP = sY + NST * tile_bytes;    // 2 x [128 x 128] bf16 (backward only)
  TcSmem* sb = reinterpret_cast<TcSmem*>(sP + (BWD ? 2 * 2 * ATOM_BYTES : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    mbar_init(&sb->d_full, 1);
    mbar_init(&sb->d_empty, TC_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(&sb->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  const uint32_t tmem_s[TC_SBUF] = {tmem, tmem + TCN, tmem + 2 * TCN};
  const uint32_t tmem_d = tmem + TC_SBUF * TCN;

  if (warp > TC_MMA_WARP) {
    // =========================== operand loaders ===========================
    const int lt = threadIdx.x - (TC_MMA_WARP + 1) * 32;
    uint32_t x_iter = 0, y_iter = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++x_iter) {
      const int rb = item / p.n_chunks, ch = item - rb * p.n_chunks;
      const int t0 = ch * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.n_col_tiles);
      const int xb = p.n_xbuf == 2 ? (x_iter & 1) : 0;
      unsigned char* sXb = sX + xb * tile_bytes;
      mbar_wait(&sb->x_empty[xb], ((p.n_xbuf == 2 ? (x_iter >> 1) : x_iter) & 1) ^ 1);
      if (p.use_tma) {
        if (lt == 0) {
          mbar_arrive_expect_tx(&sb->x_full[xb], (uint32_t)tile_bytes);
          for (int a = 0; a < NA; ++a) tma_load_2d(sXb + a * ATOM_BYTES, &map_x, a * 64, rb * TCM, &sb->x_full[xb]);
        }
      } else {
        manual_tile_load(sXb, p.X, (int64_t)rb * TCM, p.Bx, p.F, lt);
        fence_proxy_async();
        mbar_arrive(&sb->x_full[xb]);
      }
      for (int t = t0; t < t1; ++t, ++y_iter) {
        const int st = y_iter % NST;
        const uint32_t ph = (y_iter / NST) & 1;
        if (p.dbg_ts && blockIdx.x == 0 && lt == 0 && y_iter < 500) p.dbg_ts[2 * 2048 + y_iter * 4 + 0] = clock64();
        mbar_wait(&sb->y_empty[st], ph ^ 1);
        if (p.dbg_ts && blockIdx.x == 0 && lt == 0 && y_iter < 500) p.dbg_ts[2 * 2048 + y_iter * 4 + 1] = clock64();
        unsigned char* dst = sY + st * tile_bytes;
        if (p.use_tma) {
          if (lt == 0) {
            if ((p.debug & 8) && y_iter >= (uint32_t)NST) {  // profiling knob: reuse stale stage data, no TMA traffic
              mbar_arrive(&sb->y_full[st]);
            } else {
              mbar_arrive_expect_tx(&sb->y_full[st], (uint32_t)tile_bytes);
              for (int a = 0; a < NA; ++a) tma_load_2d(dst + a * ATOM_BYTES, &map_y, a * 64, t * TCN, &sb->y_full[st]);
            }
          }
        } else {
          manual_tile_load(dst, p.Y, (int64_t)t * TCN, p.By, p.F, lt);
          fence_proxy_async();
          mbar_arrive(&sb->y_full[st]);
        }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // =========================== MMA issuer (one lane) ===========================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(TCM, TCN, 0, 0);
      const uint32_t idesc2 = make_idesc(TCM, p.F, 0, 1);  // dX += P(K-major) * Y(MN-major)
      const uint32_t sX_a = smem_u32(sX), sY_a = smem_u32(sY), sP_a = smem_u32(sP);
      const int ksteps1 = p.F >> 4;
      uint32_t x_iter = 0, y_iter = 0, s_iter = 0, p_iter = 0, d_phase = 0;
      // descriptors differ only in their 14-bit start-address field: build them once, then one 64-bit add per operand
      const uint64_t dX0 = make_desc(sX_a, 16, 1024), dY0 = make_desc(sY_a, 16, 1024);
      const uint64_t dP0 = make_desc(sP_a, 16, 1024), dYmn0 = make_desc(sY_a, ATOM_BYTES, 1024);
      const uint32_t tile16 = (uint32_t)tile_bytes >> 4;
      uint64_t dXb = dX0;
      auto issue_mma1 = [&](uint32_t yi, uint32_t si) {
        const int st = yi % NST;
        const int sbuf = si % TC_SBUF;
        const bool ts = p.dbg_ts && blockIdx.x == 0 && si < 500;
        if (ts) p.dbg_ts[0 * 2048 + si * 4 + 0] = clock64();
        mbar_wait(&sb->y_full[st], (yi / NST) & 1);
        if (ts) p.dbg_ts[0 * 2048 + si * 4 + 1] = clock64();
        mbar_wait(&sb->s_empty[sbuf], ((si / TC_SBUF) & 1) ^ 1);
        if (ts) p.dbg_ts[0 * 2048 + si * 4 + 2] = clock64();
        tc_fence_after();
        const uint64_t dYs = dY0 + (uint64_t)(st * tile16);
        const uint32_t d_tmem = tmem_s[sbuf];
        if (!(p.debug & 16)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, dXb + 2 * k, dYs + 2 * k, idesc1, k > 0 ? 1u : 0u);
          if (ksteps1 > 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, dXb + (ATOM_BYTES >> 4) + 2 * k, dYs + (ATOM_BYTES >> 4) + 2 * k, idesc1, 1u);
          }
        }
        umma_commit(&sb->s_full[sbuf]);
        if (!BWD) umma_commit(&sb->y_empty[st]);
        if (ts) p.dbg_ts[0 * 2048 + si * 4 + 3] = clock64();
      };
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++x_iter) {
        const int ch = item % p.n_chunks;
        const int t0 = ch * p.tiles_per_chunk;
        const int nt = min(t0 + p.tiles_per_chunk, p.n_col_tiles) - t0;
        const int xb = p.n_xbuf == 2 ? (x_iter & 1) : 0;
        dXb = dX0 + (uint64_t)(xb * tile16);
        mbar_wait(&sb->x_full[xb], (p.n_xbuf == 2 ? (x_iter >> 1) : x_iter) & 1);
        if (!BWD) {
          for (int t = 0; t < nt; ++t) issue_mma1(y_iter++, s_iter++);
          umma_commit(&sb->x_empty[xb]);
        } else {
          bool x_released = false;
          const uint32_t y_base = y_iter;
          issue_mma1(y_iter++, s_iter++);
          if (nt > 1) issue_mma1(y_iter++, s_iter++);
          for (int t = 0; t < nt; ++t) {
            if (t + 2 < nt) issue_mma1(y_iter++, s_iter++);
            if (t + 2 >= nt && !x_released) {  // every MMA that reads X of this item has been issued
              umma_commit(&sb->x_empty[xb]);
              x_released = true;
            }
            const int pb = p_iter & 1;
            mbar_wait(&sb->p_full[pb], (p_iter >> 1) & 1);
            if (t == 0) mbar_wait(&sb->d_empty, d_phase ^ 1);
            tc_fence_after();
            const int st = (y_base + t) % NST;
            const uint64_t dPb = dP0 + (uint64_t)(pb * ((2 * ATOM_BYTES) >> 4));
            const uint64_t dYm = dYmn0 + (uint64_t)(st * tile16);
#pragma unroll
            for (int kk = 0; kk < TCN / 16; ++kk)
              umma_bf16(tmem_d, dPb + (uint64_t)((kk >> 2) * (ATOM_BYTES >> 4) + (kk & 3) * 2), dYm + (uint64_t)(kk * (2048 >> 4)), idesc2,
                        (t > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&sb->p_empty[pb]);
            umma_commit(&sb->y_empty[st]);
            ++p_iter;
          }
          umma_commit(&sb->d_full);
          d_phase ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue warps 0-7 ===========================
    // thread == one row of the block (TMEM lane quarter = warp % 4) and one 32-column slice of every tile (warp / 4)
    const int quarter = warp & 3, slice = warp >> 2;
    const int r_in_blk = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int et = threadIdx.x;  // 0..511
    float* s_iy = sb->iy;        // [TC_SBUF][128] gscale * inv_y of the current tiles (backward)
    uint32_t s_iter = 0, p_iter = 0, d_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int rb = item / p.n_chunks, ch = item - rb * p.n_chunks;
      const int t0 = ch * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.n_col_tiles);
      const int64_t row = (int64_t)rb * TCM + r_in_blk;
      const bool row_ok = row < p.Bx;
      const bool rows_full = (int64_t)rb * TCM + TCM <= p.Bx;
      const int64_t dcol = p.diag_offset + row;  // this row's positive column
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, accs[4] = {0.f, 0.f, 0.f, 0.f};
      float best = -FLT_MAX, dval = 0.f;
      int32_t best_j = 0x7fffffff;
      bool have_diag = false;
      float ixs = 0.f, iy_next = 0.f;
      if (BWD) {
        ixs = row_ok ? p.gscale * p.inv_x[row] : 0.f;
        if (et < TCN) {
          const int64_t c = (int64_t)t0 * TCN + et;
          iy_next = c < p.By ? p.gscale * p.inv_y[c] : 0.f;
        }
      }
      for (int t = t0; t < t1; ++t, ++s_iter) {
        const int sbuf = s_iter % TC_SBUF;
        const int64_t col0 = (int64_t)t * TCN;
        const bool partial = col0 + TCN > p.By;
        const bool diag_tile = (p.diag_offset + (int64_t)rb * TCM < col0 + TCN) && (p.diag_offset + (int64_t)rb * TCM + TCM > col0);
        const bool generic = partial || diag_tile || (BWD && !rows_full);
        int pb = 0;
        if (BWD) {
          // publish this tile's column scales, prefetch the next tile's
          if (et < TCN) s_iy[sbuf * TCN + et] = iy_next;
          asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
          if (et < TCN && t + 1 < t1) {
            const int64_t c = col0 + TCN + et;
            iy_next = c < p.By ? p.gscale * p.inv_y[c] : 0.f;
          }
          pb = p_iter & 1;
          mbar_wait(&sb->p_empty[pb], ((p_iter >> 1) & 1) ^ 1);
        }
        const bool stamp = p.dbg_ts && blockIdx.x == 0 && threadIdx.x == 0 && s_iter < 500;
        if (stamp) p.dbg_ts[1 * 2048 + s_iter * 4 + 0] = clock64();
        mbar_wait(&sb->s_full[sbuf], (s_iter / TC_SBUF) & 1);
        if (stamp) p.dbg_ts[1 * 2048 + s_iter * 4 + 1] = clock64();
        tc_fence_after();
        uint32_t r0[32];
        if (!(p.debug & 2)) {
          tmem_ld32(tmem_s[sbuf] + lane_addr + slice * 32, r0);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r0[j] = 0x3c000000u + j;
        }
        const int64_t cb = col0 + slice * 32;
        if (!BWD) {
          if (generic) fwd_chunk<true, true>(r0, cb, p, dcol, acc, accs, best, best_j, dval, have_diag);
          else if (p.want_sum_s) fwd_chunk<false, true>(r0, cb, p, dcol, acc, accs, best, best_j, dval, have_diag);
          else fwd_chunk<false, false>(r0, cb, p, dcol, acc, accs, best, best_j, dval, have_diag);
        } else {
          unsigned char* prow = sP + pb * 2 * ATOM_BYTES + (slice >> 1) * ATOM_BYTES + r_in_blk * 128;
          const float* iy = s_iy + sbuf * TCN;
          if (generic) bwd_chunk<true>(r0, cb, slice * 32, p, dcol, row_ok, ixs, iy, prow, r_in_blk & 7);
          else bwd_chunk<false>(r0, cb, slice * 32, p, dcol, row_ok, ixs, iy, prow, r_in_blk & 7);
          fence_proxy_async();
        }
        if (stamp) p.dbg_ts[1 * 2048 + s_iter * 4 + 2] = clock64();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sb->s_empty[sbuf]);
          if (BWD) mbar_arrive(&sb->p_full[pb]);
        }
        if (stamp) p.dbg_ts[1 * 2048 + s_iter * 4 + 3] = clock64();
        if (BWD) ++p_iter;
      }
      if (!BWD) {
        if (row_ok) {
          const int64_t o = ((int64_t)ch * 4 + slice) * p.Bx + row;
          p.row_e[o] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
          if (p.want_sum_s) p.row_s[o] = ((accs[0] + accs[1]) + (accs[2] + accs[3])) * p.inv_temp;
          if (p.want_metrics) {
            p.row_max[o] = best * p.inv_temp;
            p.row_arg[o] = best_j;
            if (have_diag) p.diag[row] = dval * p.inv_temp;
          }
        }
      } else {
        mbar_wait(&sb->d_full, d_phase);
        d_phase ^= 1;
        tc_fence_after();
        // 32-column slices of the F output columns: slice s < F/32 is read out by the warps with warp / 4 == s
        if (slice * 32 < p.F) {
          float* out = p.dx_part + ((int64_t)ch * p.Bx + row) * p.F + slice * 32;
          uint32_t r[32];
          tmem_ld32(tmem_d + lane_addr + slice * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(out + q * 4) =
                  make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                              __uint_as_float(r[q * 4 + 3]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sb->d_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------- small deterministic reductions
__global__ void tc_rows_reduce_kernel(const float* __restrict__ row_e, const float* __restrict__ row_s,
                                      const float* __restrict__ row_max, const int32_t* __restrict__ row_arg, int64_t Bx,
                                      int n_parts, float* __restrict__ out_e, float* __restrict__ out_s,
                                      int32_t* __restrict__ out_arg) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Bx) return;
  float e = 0.f, s = 0.f, mx = -FLT_MAX;
  int32_t arg = 0x7fffffff;
#pragma unroll 8
  for (int c = 0; c < n_parts; ++c) {
    const int64_t o = (int64_t)c * Bx + m;
    e += row_e[o];
    if (out_s) s += row_s[o];
    if (out_arg) {
      const float tm = row_max[o];
      const int32_t ta = row_arg[o];
      if (tm > mx || (tm == mx && ta < arg)) {
        mx = tm;
        arg = ta;
      }
    }
  }
  out_e[m] = e;
  if (out_s) out_s[m] = s;
  if (out_arg) out_arg[m] = arg;
}

__global__ void tc_dx_reduce_kernel(const float* __restrict__ part, int64_t n, int n_chunks, float* __restrict__ dX) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 a = *reinterpret_cast<const float4*>(part + i);
  for (int c = 1; c < n_chunks; ++c) f4_add(a, *reinterpret_cast<const float4*>(part + (int64_t)c * n + i));
  *reinterpret_cast<float4*>(dX + i) = a;
}

// ---------------------------------------------------------------- host side
static bool make_map(CUtensorMap* map, const void* ptr, int64_t rows, int F) {
  return make_map_bf16(map, ptr, rows, F, F, 64, TCM);
}

struct TcPlan {
  int n_row_blocks, n_col_tiles, tiles_per_chunk, n_chunks, n_items, grid;
};

// split the column tiles so that (row blocks x chunks) fills whole waves of the machine, >= 4 tiles per chunk if possible
static TcPlan make_plan(int64_t Bx, int64_t By) {
  TcPlan pl;
  pl.n_row_blocks = (int)((Bx + TCM - 1) / TCM);
  pl.n_col_tiles = (int)((By + TCN - 1) / TCN);
  const int sms = sm_count();
  int best_s = 1;
  double best_eff = -1.0;
  const int max_s = pl.n_col_tiles < 64 ? pl.n_col_tiles : 64;
  for (int s = 1; s <= max_s; ++s) {
    const int tpc = (pl.n_col_tiles + s - 1) / s;
    if (tpc < 4 && s > 1) break;
    const int chunks = (pl.n_col_tiles + tpc - 1) / tpc;
    const long items = (long)pl.n_row_blocks * chunks;
    const long waves = (items + sms - 1) / sms;
    const double eff = (double)items / (double)(waves * sms) - 0.002 * s;  // mild preference for fewer partials
    if (eff > best_eff) {
      best_eff = eff;
      best_s = s;
    }
  }
  pl.tiles_per_chunk = (pl.n_col_tiles + best_s - 1) / best_s;
  pl.n_chunks = (pl.n_col_tiles + pl.tiles_per_chunk - 1) / pl.tiles_per_chunk;
  pl.n_items = pl.n_row_blocks * pl.n_chunks;
  pl.grid = pl.n_items < sms ? pl.n_items : sms;
  return pl;
}

static void tc_buffers(int F, bool bwd, int& n_stages, int& n_xbuf) {
  const bool tight = bwd && F > 64;  // 32 KB tiles + 64 KB of dS buffers: stay under 227 KB
  n_stages = tight ? 3 : TC_STAGES;
  n_xbuf = tight ? 1 : 2;
}
static size_t tc_smem_bytes(int F, bool bwd) {
  const int tile = (F >> 6) * ATOM_BYTES;
  int ns, nx;
  tc_buffers(F, bwd, ns, nx);
  return 1024 + (size_t)tile * (nx + ns) + (bwd ? 4 * ATOM_BYTES : 0) + sizeof(TcSmem) + 64;
}

int64_t simce_tc_workspace_bytes(int64_t Bx, int64_t By, int F) {
  const int64_t B = Bx > By ? Bx : By;
  TcPlan a = make_plan(Bx, By), b = make_plan(By, Bx);
  const int64_t chunks = a.n_chunks > b.n_chunks ? a.n_chunks : b.n_chunks;
  const int64_t fwd = 4 * ((4 * chunks * B * 4 + 255) / 256 * 256) + 256;
  const int64_t bwd = chunks * B * F * 4 + 256;
  return (fwd > bwd ? fwd : bwd) + 1024;
}

static int g_tc_force_manual = 0;
static int g_tc_debug = 0;
static long long* g_tc_dbg_ts = nullptr;

// rows statistics of S = X Y^T (bf16 inputs).  want_metrics: also max/argmax/diag.
int simce_tc_rows(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                  float* row_sumexp, float* row_sum_s, float* diag, int32_t* row_argmax, int want_metrics, void* workspace,
                  cudaStream_t s) {
  TcPlan pl = make_plan(Bx, By);
  TcParams p = {};
  p.X = (const __nv_bfloat16*)X;
  p.Y = (const __nv_bfloat16*)Y;
  p.Bx = Bx; p.By = By; p.F = F;
  p.n_col_tiles = pl.n_col_tiles; p.tiles_per_chunk = pl.tiles_per_chunk; p.n_chunks = pl.n_chunks; p.n_items = pl.n_items;
  const float log2e = 1.4426950408889634f;
  p.c1 = inv_temp * log2e; p.c0 = -inv_temp * log2e; p.inv_temp = inv_temp;
  p.diag_offset = diag_offset;
  const int64_t stride = (4 * pl.n_chunks * Bx * 4 + 255) / 256 * 256;
  char* w = (char*)workspace;
  p.row_e = (float*)w; p.row_s = (float*)(w + 3 * stride); p.row_max = (float*)(w + stride); p.row_arg = (int32_t*)(w + 2 * stride);
  p.want_sum_s = row_sum_s != nullptr;
  p.diag = diag; p.want_metrics = want_metrics; p.debug = g_tc_debug; p.dbg_ts = want_metrics ? g_tc_dbg_ts : nullptr;
  tc_buffers(F, false, p.n_stages, p.n_xbuf);
  CUtensorMap mx, my;
  p.use_tma = (!g_tc_force_manual && make_map(&mx, X, Bx, F) && make_map(&my, Y, By, F)) ? 1 : 0;
  if (!p.use_tma) { memset(&mx, 0, sizeof(mx)); memset(&my, 0, sizeof(my)); }
  const size_t smem = tc_smem_bytes(F, false);
  TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simce_tc_kernel<false><<<pl.grid, TC_THREADS, smem, s>>>(mx, my, p);
  tc_rows_reduce_kernel<<<ceil_div(Bx, 128), 128, 0, s>>>(p.row_e, p.row_s, p.row_max, p.row_arg, Bx, 4 * pl.n_chunks, row_sumexp,
                                                          row_sum_s, want_metrics ? row_argmax : nullptr);
  TT_LAUNCH_CHECK("simce_tc_rows");
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
  p.n_col_tiles = pl.n_col_tiles; p.tiles_per_chunk = pl.tiles_per_chunk; p.n_chunks = pl.n_chunks; p.n_items = pl.n_items;
  const float log2e = 1.4426950408889634f;
  p.c1 = inv_temp * log2e; p.c0 = -inv_temp * log2e; p.inv_temp = inv_temp;
  p.diag_offset = diag_offset;
  p.inv_x = inv_x; p.inv_y = inv_y; p.gscale = gscale;
  p.tgt_diag2 = 2.f * gscale * tgt_diag; p.tgt_off2 = 2.f * gscale * tgt_off;
  p.dx_part = pl.n_chunks == 1 ? dX : (float*)workspace;
  tc_buffers(F, true, p.n_stages, p.n_xbuf);
  CUtensorMap mx, my;
  p.use_tma = (!g_tc_force_manual && make_map(&mx, X, Bx, F) && make_map(&my, Y, By, F)) ? 1 : 0;
  if (!p.use_tma) { memset(&mx, 0, sizeof(mx)); memset(&my, 0, sizeof(my)); }
  const size_t smem = tc_smem_bytes(F, true);
  TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simce_tc_kernel<true><<<pl.grid, TC_THREADS, smem, s>>>(mx, my, p);
  if (pl.n_chunks > 1)
    tc_dx_reduce_kernel<<<ceil_div(Bx * F / 4, 256), 256, 0, s>>>(p.dx_part, Bx * F, pl.n_chunks, dX);
  TT_LAUNCH_CHECK("simce_tc_bwd");
  return TT_OK;
}

void simce_tc_set_force_manual(int v) { g_tc_force_manual = v; }
void simce_tc_set_debug(int v) { g_tc_debug = v; }
void simce_tc_set_timeline(long long* buf) { g_tc_dbg_ts = buf; }

}  // namespace tt
