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
// Warp roles (224 threads): warps 0-3 epilogue (TMEM lane quarter == warp id), warp 4 MMA issuer + TMEM
// allocator, warps 5-6 operand loaders (TMA: one elected lane issues cp.async.bulk.tensor with 128B swizzle;
// the generic-proxy fallback loads 16-byte chunks and applies the same swizzle by hand).
// Pipelines: Y tiles through a 3-stage smem ring (full/empty mbarriers), S through 2 TMEM buffers, dS through
// 2 smem buffers; tcgen05.commit arrives on the barriers when the MMAs that read a buffer have retired.
#include "common.cuh"
#include <cuda.h>
#include <float.h>

namespace tt {

constexpr int TCM = 128;          // rows per CTA block (UMMA M)
constexpr int TCN = 128;          // columns per tile  (UMMA N)
constexpr int TC_STAGES = 3;      // Y smem ring
constexpr int TC_THREADS = 224;   // 4 epilogue + 1 mma + 2 loader warps
constexpr int TC_EPI_THREADS = 128;
constexpr int TC_LOADER_THREADS = 64;
constexpr int ATOM_BYTES = TCM * 128;  // one 128-row x 64-bf16 swizzle atom = 16 KB

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// bounded spin: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (done) return;
  }
  printf("simce_tc: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// smem matrix descriptor, SWIZZLE_128B, version 1 (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, M x N, majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

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
  // forward outputs: partials per (chunk, row)
  float* row_e;
  float* row_s;
  float* row_max;
  int32_t* row_arg;
  float* diag;
  int want_metrics;
  // backward
  const float* inv_x;    // [Bx] (unscaled)
  const float* inv_y;    // [By]
  float gscale, tgt_diag2, tgt_off2;  // 2*gscale*target values
  float* dx_part;        // [n_chunks][Bx][F]
};

struct TcSmem {  // carved from dynamic smem after the 1024-aligned tiles
  uint64_t x_full, x_empty;
  uint64_t y_full[TC_STAGES], y_empty[TC_STAGES];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full[2], p_empty[2];
  uint64_t d_full, d_empty;
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

template <bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
simce_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NA = p.F >> 6;                      // swizzle atoms along K (1 or 2)
  const int tile_bytes = NA * ATOM_BYTES;       // X / Y tile
  unsigned char* sX = base;
  unsigned char* sY = sX + tile_bytes;          // TC_STAGES tiles
  unsigned char* sP = sY + TC_STAGES * tile_bytes;  // 2 x [128 x 128] bf16 (backward only)
  TcSmem* sb = reinterpret_cast<TcSmem*>(sP + (BWD ? 2 * 2 * ATOM_BYTES : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_loader_arrivals = p.use_tma ? 1 : TC_LOADER_THREADS;

  if (threadIdx.x == 0) {
    mbar_init(&sb->x_full, n_loader_arrivals);
    mbar_init(&sb->x_empty, 1);
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&sb->y_full[i], n_loader_arrivals);
      mbar_init(&sb->y_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb->s_full[i], 1);
      mbar_init(&sb->s_empty[i], TC_EPI_THREADS);
      mbar_init(&sb->p_full[i], TC_EPI_THREADS);
      mbar_init(&sb->p_empty[i], 1);
    }
    mbar_init(&sb->d_full, 1);
    mbar_init(&sb->d_empty, TC_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&sb->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sb->tmem_base;
  const uint32_t tmem_s[2] = {tmem, tmem + TCN};
  const uint32_t tmem_d = tmem + 2 * TCN;

  const int n_row_blocks = (int)((p.Bx + TCM - 1) / TCM);
  (void)n_row_blocks;

  if (warp >= 5) {
    // =========================== operand loaders ===========================
    const int lt = threadIdx.x - 5 * 32;
    uint32_t x_phase = 0, y_iter = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int rb = item / p.n_chunks, ch = item - rb * p.n_chunks;
      const int t0 = ch * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.n_col_tiles);
      mbar_wait(&sb->x_empty, x_phase ^ 1);
      if (p.use_tma) {
        if (lt == 0) {
          mbar_arrive_expect_tx(&sb->x_full, (uint32_t)tile_bytes);
          for (int a = 0; a < NA; ++a) tma_load_2d(sX + a * ATOM_BYTES, &map_x, a * 64, rb * TCM, &sb->x_full);
        }
      } else {
        manual_tile_load(sX, p.X, (int64_t)rb * TCM, p.Bx, p.F, lt);
        fence_proxy_async();
        mbar_arrive(&sb->x_full);
      }
      x_phase ^= 1;
      for (int t = t0; t < t1; ++t, ++y_iter) {
        const int st = y_iter % TC_STAGES;
        const uint32_t ph = (y_iter / TC_STAGES) & 1;
        mbar_wait(&sb->y_empty[st], ph ^ 1);
        unsigned char* dst = sY + st * tile_bytes;
        if (p.use_tma) {
          if (lt == 0) {
            mbar_arrive_expect_tx(&sb->y_full[st], (uint32_t)tile_bytes);
            for (int a = 0; a < NA; ++a) tma_load_2d(dst + a * ATOM_BYTES, &map_y, a * 64, t * TCN, &sb->y_full[st]);
          }
        } else {
          manual_tile_load(dst, p.Y, (int64_t)t * TCN, p.By, p.F, lt);
          fence_proxy_async();
          mbar_arrive(&sb->y_full[st]);
        }
      }
    }
  } else if (warp == 4) {
    // =========================== MMA issuer (one lane) ===========================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(TCM, TCN, 0, 0);
      const uint32_t idesc2 = make_idesc(TCM, p.F, 0, 1);  // dX += P(K-major) * Y(MN-major)
      const uint32_t sX_a = smem_u32(sX), sY_a = smem_u32(sY), sP_a = smem_u32(sP);
      const int ksteps1 = p.F >> 4;
      uint32_t x_phase = 0, y_iter = 0, s_iter = 0, p_iter = 0, d_phase = 0;
      auto issue_mma1 = [&](uint32_t yi, uint32_t si) {
        const int st = yi % TC_STAGES;
        const int sbuf = si & 1;
        mbar_wait(&sb->y_full[st], (yi / TC_STAGES) & 1);
        mbar_wait(&sb->s_empty[sbuf], ((si >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int k = 0; k < ksteps1; ++k) {
          const uint32_t off = (uint32_t)((k >> 2) * ATOM_BYTES + (k & 3) * 32);
          umma_bf16(tmem_s[sbuf], make_desc(sX_a + off, 16, 1024), make_desc(sY_a + st * tile_bytes + off, 16, 1024), idesc1,
                    k > 0 ? 1u : 0u);
        }
        umma_commit(&sb->s_full[sbuf]);
        if (!BWD) umma_commit(&sb->y_empty[st]);
      };
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ch = item % p.n_chunks;
        const int t0 = ch * p.tiles_per_chunk;
        const int nt = min(t0 + p.tiles_per_chunk, p.n_col_tiles) - t0;
        mbar_wait(&sb->x_full, x_phase);
        x_phase ^= 1;
        if (!BWD) {
          for (int t = 0; t < nt; ++t) issue_mma1(y_iter++, s_iter++);
          umma_commit(&sb->x_empty);
        } else {
          const uint32_t y_base = y_iter;
          issue_mma1(y_iter++, s_iter++);
          for (int t = 0; t < nt; ++t) {
            if (t + 1 < nt) issue_mma1(y_iter++, s_iter++);
            if (t + 1 == nt) umma_commit(&sb->x_empty);  // every MMA that reads X of this item has been issued
            const int pb = p_iter & 1;
            mbar_wait(&sb->p_full[pb], (p_iter >> 1) & 1);
            if (t == 0) mbar_wait(&sb->d_empty, d_phase ^ 1);
            tc_fence_after();
            const int st = (y_base + t) % TC_STAGES;
            for (int kk = 0; kk < TCN / 16; ++kk) {
              const uint64_t a_desc = make_desc(sP_a + pb * 2 * ATOM_BYTES + (kk >> 2) * ATOM_BYTES + (kk & 3) * 32, 16, 1024);
              const uint64_t b_desc = make_desc(sY_a + st * tile_bytes + kk * 16 * 128, ATOM_BYTES, 1024);
              umma_bf16(tmem_d, a_desc, b_desc, idesc2, (t > 0 || kk > 0) ? 1u : 0u);
            }
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
    // =========================== epilogue warps 0-3: thread == one row of the block ===========================
    const int r_in_blk = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t s_iter = 0, p_iter = 0, d_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int rb = item / p.n_chunks, ch = item - rb * p.n_chunks;
      const int t0 = ch * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.n_col_tiles);
      const int64_t row = (int64_t)rb * TCM + r_in_blk;
      const bool row_ok = row < p.Bx;
      const int64_t dcol = p.diag_offset + row;  // this row's positive column
      float acc_e = 0.f, acc_s = 0.f, best = -FLT_MAX, dval = 0.f;
      int32_t best_j = 0x7fffffff;
      bool have_diag = false;
      float ixs = 0.f;
      if (BWD) ixs = row_ok ? p.gscale * p.inv_x[row] : 0.f;
      for (int t = t0; t < t1; ++t, ++s_iter) {
        const int sbuf = s_iter & 1;
        const int64_t col0 = (int64_t)t * TCN;
        const bool partial = col0 + TCN > p.By;
        const bool diag_tile = (p.diag_offset + (int64_t)rb * TCM < col0 + TCN) && (p.diag_offset + (int64_t)rb * TCM + TCM > col0);
        mbar_wait(&sb->s_full[sbuf], (s_iter >> 1) & 1);
        tc_fence_after();
        int pb = 0;
        if (BWD) {
          pb = p_iter & 1;
          mbar_wait(&sb->p_empty[pb], ((p_iter >> 1) & 1) ^ 1);
        }
#pragma unroll 1
        for (int c32 = 0; c32 < TCN / 32; ++c32) {
          uint32_t r[32];
          tmem_ld32(tmem_s[sbuf] + lane_addr + c32 * 32, r);
          tmem_ld_wait();
          const int64_t cbase = col0 + c32 * 32;
          if (!BWD) {
            if (!partial && !diag_tile) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float raw = __uint_as_float(r[j]);
                acc_e += ex2f(fmaf(raw, p.c1, p.c0));
                acc_s += raw;
                if (raw > best) {
                  best = raw;
                  best_j = (int32_t)(cbase + j);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float raw = __uint_as_float(r[j]);
                const int64_t col = cbase + j;
                if (col < p.By) {
                  acc_e += ex2f(fmaf(raw, p.c1, p.c0));
                  acc_s += raw;
                  if (raw > best) {
                    best = raw;
                    best_j = (int32_t)col;
                  }
                  if (col == dcol) {
                    dval = raw;
                    have_diag = true;
                  }
                }
              }
            }
          } else {
            uint32_t packed[16];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 iy = make_float4(0.f, 0.f, 0.f, 0.f);
              if (!partial) iy = __ldg(reinterpret_cast<const float4*>(p.inv_y + cbase) + j4);
              else {
                if (cbase + j4 * 4 + 0 < p.By) iy.x = p.inv_y[cbase + j4 * 4 + 0];
                if (cbase + j4 * 4 + 1 < p.By) iy.y = p.inv_y[cbase + j4 * 4 + 1];
                if (cbase + j4 * 4 + 2 < p.By) iy.z = p.inv_y[cbase + j4 * 4 + 2];
                if (cbase + j4 * 4 + 3 < p.By) iy.w = p.inv_y[cbase + j4 * 4 + 3];
              }
              const float iyv[4] = {iy.x, iy.y, iy.z, iy.w};
              float d[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int j = j4 * 4 + q;
                const float raw = __uint_as_float(r[j]);
                const float e = ex2f(fmaf(raw, p.c1, p.c0));
                float v = fmaf(e, fmaf(p.gscale, iyv[q], ixs), -p.tgt_off2);
                if (diag_tile && (cbase + j == dcol)) v = fmaf(e, fmaf(p.gscale, iyv[q], ixs), -p.tgt_diag2);
                if (partial && (cbase + j >= p.By)) v = 0.f;
                d[q] = row_ok ? v : 0.f;
              }
              __nv_bfloat162 lo = __floats2bfloat162_rn(d[0], d[1]);
              __nv_bfloat162 hi = __floats2bfloat162_rn(d[2], d[3]);
              packed[j4 * 2 + 0] = *reinterpret_cast<uint32_t*>(&lo);
              packed[j4 * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
            }
            // 32 columns = four 16-byte chunks of this row inside atom (c32 / 2)
            unsigned char* prow = sP + pb * 2 * ATOM_BYTES + (c32 >> 1) * ATOM_BYTES + r_in_blk * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int cc = (c32 & 1) * 4 + q;
              *reinterpret_cast<uint4*>(prow + ((cc ^ (r_in_blk & 7)) << 4)) =
                  make_uint4(packed[q * 4 + 0], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&sb->s_empty[sbuf]);
        if (BWD) {
          fence_proxy_async();
          mbar_arrive(&sb->p_full[pb]);
          ++p_iter;
        }
      }
      if (!BWD) {
        if (row_ok) {
          const int64_t o = (int64_t)ch * p.Bx + row;
          p.row_e[o] = acc_e;
          p.row_s[o] = acc_s * p.inv_temp;
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
        float* out = p.dx_part + ((int64_t)ch * p.Bx + row) * p.F;
        for (int c32 = 0; c32 < (p.F >> 5); ++c32) {
          uint32_t r[32];
          tmem_ld32(tmem_d + lane_addr + c32 * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(out + c32 * 32 + q * 4) =
                  make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                              __uint_as_float(r[q * 4 + 3]));
          }
        }
        tc_fence_before();
        mbar_arrive(&sb->d_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------- small deterministic reductions
__global__ void tc_rows_reduce_kernel(const float* __restrict__ row_e, const float* __restrict__ row_s,
                                      const float* __restrict__ row_max, const int32_t* __restrict__ row_arg, int64_t Bx,
                                      int n_chunks, float* __restrict__ out_e, float* __restrict__ out_s,
                                      int32_t* __restrict__ out_arg) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Bx) return;
  float e = 0.f, s = 0.f, mx = -FLT_MAX;
  int32_t arg = 0x7fffffff;
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t o = (int64_t)c * Bx + m;
    e += row_e[o];
    s += row_s[o];
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

// [rows, F] bf16 row-major; box = 64 columns (128 B, swizzle span) x 128 rows
static bool make_map(CUtensorMap* map, const void* ptr, int64_t rows, int F) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)F * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)TCM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
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

static size_t tc_smem_bytes(int F, bool bwd) {
  const int tile = (F >> 6) * ATOM_BYTES;
  return 1024 + (size_t)tile * (1 + TC_STAGES) + (bwd ? 4 * ATOM_BYTES : 0) + sizeof(TcSmem) + 64;
}

int64_t simce_tc_workspace_bytes(int64_t Bx, int64_t By, int F) {
  const int64_t B = Bx > By ? Bx : By;
  TcPlan a = make_plan(Bx, By), b = make_plan(By, Bx);
  const int64_t chunks = a.n_chunks > b.n_chunks ? a.n_chunks : b.n_chunks;
  const int64_t fwd = 4 * ((chunks * B * 4 + 255) / 256 * 256);
  const int64_t bwd = chunks * B * F * 4 + 256;
  return (fwd > bwd ? fwd : bwd) + 1024;
}

static int g_tc_force_manual = 0;

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
  const int64_t stride = (pl.n_chunks * Bx * 4 + 255) / 256 * 256;
  char* w = (char*)workspace;
  p.row_e = (float*)w; p.row_s = (float*)(w + stride); p.row_max = (float*)(w + 2 * stride); p.row_arg = (int32_t*)(w + 3 * stride);
  p.diag = diag; p.want_metrics = want_metrics;
  CUtensorMap mx, my;
  p.use_tma = (!g_tc_force_manual && make_map(&mx, X, Bx, F) && make_map(&my, Y, By, F)) ? 1 : 0;
  if (!p.use_tma) { memset(&mx, 0, sizeof(mx)); memset(&my, 0, sizeof(my)); }
  const size_t smem = tc_smem_bytes(F, false);
  TT_CUDA(cudaFuncSetAttribute(simce_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simce_tc_kernel<false><<<pl.grid, TC_THREADS, smem, s>>>(mx, my, p);
  tc_rows_reduce_kernel<<<ceil_div(Bx, 256), 256, 0, s>>>(p.row_e, p.row_s, p.row_max, p.row_arg, Bx, pl.n_chunks, row_sumexp,
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

}  // namespace tt
