// Embedding-bag collection: pooled lookup, deterministic sort-by-row dedup, segment-sum + fused Adam.
//
// HBM-bound byte work: every kernel here is a coalesced 128-bit gather/scatter.  One "lane group"
// of E/4 lanes owns one embedding row (E=32 fp32 -> 8 lanes x float4 = one 128-byte line), so a
// warp moves 4 rows per instruction and consecutive bags of a sample land contiguously in the
// tower's concat buffer.  Table metadata (row base, vocab) is staged through shared memory.
#include "common.cuh"
#include "adam.cuh"

namespace tt {

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ void store_vec4(OutT* p, const float4& v);
template <>
__device__ __forceinline__ void store_vec4<float>(float* p, const float4& v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store_vec4<__nv_bfloat16>(__nv_bfloat16* p, const float4& v) {
  *reinterpret_cast<uint2*>(p) = pack_bf16x4(v);
}

// L = 1 fast path (the reference's only case): UNROLL bags in flight per lane group.
// LAZY (lazy exact-replay Adam) is a template flag: the replay state costs ~50 registers, which capped the plain lookup at two
// blocks per SM (ncu: 91 registers, 23 % of the warp slots in use, 34 us for the notice tower).
template <typename OutT, int UNROLL, bool LAZY>
__global__ void __launch_bounds__(256, LAZY ? 2 : 4)
ebc_fwd_l1_kernel(const float* __restrict__ arena, const int64_t* __restrict__ row_base,
                  const int32_t* __restrict__ vocab, const int64_t* __restrict__ values, int B, int K, int E,
                  int layout, OutT* __restrict__ out, int64_t out_ld, int64_t out_col0, int32_t* __restrict__ rows_out,
                  int32_t* __restrict__ slots_out, int64_t slot_base, int G, const int32_t* __restrict__ large_rank,
                  int K_large, int64_t row_sub, const tt_lazy_ctx* __restrict__ lazy, int upto_off) {
  extern __shared__ unsigned char smem_raw[];
  int64_t* s_base = reinterpret_cast<int64_t*>(smem_raw);
  int32_t* s_vocab = reinterpret_cast<int32_t*>(s_base + K);
  int32_t* s_rank = s_vocab + K;
  // lazy exact-replay Adam: rows of the sorted (large) tables may be several steps behind; they are caught up on the fly
  tt_lazy_ctx L = {};
  tt_adam_hyper lh = {};
  long long upto = 0;
  if (LAZY) {
    L = *lazy;
    lh = *L.hyper;
    upto = lh.step + upto_off;
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    s_base[i] = row_base[i];
    s_vocab[i] = vocab[i];
    s_rank[i] = large_rank ? large_rank[i] : i;
  }
  __syncthreads();
  const int lane = threadIdx.x % G;
  // 32-bit slot arithmetic (the host checks B*K < 2^31): a 64-bit division by K is a long instruction sequence per bag
  const int n_bags = B * K;
  const int groups_total = gridDim.x * (blockDim.x / G);
  const int group0 = blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
  const int K_out = large_rank ? K_large : K;
  const int vec_per_row = E >> 2;
  const int step_b = groups_total / K, step_k = groups_total - step_b * K;
  for (int64_t base = group0; base < n_bags; base += (int64_t)groups_total * UNROLL) {
    // Three phases, each over all UNROLL bags, with no branch between the loads of a phase: the UNROLL id loads leave back to
    // back, then the UNROLL row loads (a bag past the end, or one of a skipped key, reads id 0 / arena row 0 and stores nothing).
    // With the per-bag branches of the first version the compiler had serialised the bags on the id-load latency.
    int bb[UNROLL], kk[UNROLL], voc[UNROLL];
    bool ok[UNROLL];
    int64_t id[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t s64 = base + (int64_t)u * groups_total;
      int b, k;
      if (u == 0) {  // one division per thread and iteration: the UNROLL slots of a lane group are groups_total apart
        b = (int)base / K;
        k = (int)base - b * K;
      } else {
        k = kk[u - 1] + step_k;
        b = bb[u - 1] + step_b;
        if (k >= K) {
          k -= K;
          ++b;
        }
      }
      bb[u] = b;
      kk[u] = k;
      voc[u] = s_vocab[k];
      ok[u] = s64 < n_bags && voc[u] > 0;  // vocab <= 0 marks a key this launch must skip (tables served by another path)
      const int64_t vi = layout == TT_LAYOUT_SAMPLE_MAJOR ? s64 : (int64_t)k * B + b;
      id[u] = __ldg(values + (ok[u] ? vi : 0));
    }
    int64_t row[UNROLL];
    long long lag_from[LAZY ? UNROLL : 1];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t hi = (int64_t)voc[u] - 1;
      const int64_t idc = id[u] < 0 ? 0 : (id[u] > hi ? hi : id[u]);  // clamp(id, 0, V_k-1)
      row[u] = ok[u] ? s_base[kk[u]] + idc : 0;
      const int rk = s_rank[kk[u]];
      if (LAZY) {
        lag_from[u] = -1;
        if (ok[u] && rk >= 0) {
          const long long last = L.row_last[row[u] - row_sub];
          if (last < upto) lag_from[u] = last + 1;
        }
      }
      if (ok[u] && lane == 0 && rk >= 0) {  // only tables on the sort-dedup path record their (compact) row
        const int64_t o = (int64_t)bb[u] * K_out + rk;
        if (rows_out) rows_out[o] = (int32_t)(row[u] - row_sub);
        if (slots_out) slots_out[o] = (int32_t)(slot_base + o);
      }
    }
#pragma unroll 1
    for (int c = lane; c < vec_per_row; c += G) {  // lanes_per_row() gives one 16-byte chunk per lane for E <= 128: a single trip
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(arena + row[u] * E) + c);
      if (LAZY) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (ok[u] && lag_from[u] >= 0) {  // read-only catch-up: the update kernel replays again and writes back
            const int64_t lo = (row[u] - row_sub) * E;
            float4 mm = *(reinterpret_cast<const float4*>(L.arena_m + lo) + c);
            float4 vv = *(reinterpret_cast<const float4*>(L.arena_v + lo) + c);
            lazy_replay_vec4(v[u], mm, vv, lag_from[u], upto, L, lh);
          }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (ok[u]) store_vec4<OutT>(out + (int64_t)bb[u] * out_ld + out_col0 + kk[u] * E + 4 * c, v[u]);
    }
  }
}

// jagged bags (L >= 0 per bag), sum pooling. HBM-bound at scale (BASELINE config 5: pooling factor 20, E=64, tables >> L2), so the
// kernel is built for memory-level parallelism: the G lanes of a group fetch G ids of the bag at once (one coalesced load, the
// clamped arena rows and gradient slots are recorded with coalesced stores), then rows are gathered JAG_MLP at a time — all
// loads of a batch are issued before the first add. The sum runs in ascending value position (embedding_bag's order).
constexpr int JAG_MLP = 8;
template <typename OutT>
__global__ void __launch_bounds__(256, 3)
ebc_fwd_jagged_kernel(const float* __restrict__ arena, const int64_t* __restrict__ row_base,
                      const int32_t* __restrict__ vocab, const int64_t* __restrict__ values,
                      const int64_t* __restrict__ bag_offsets, int B, int K, int E, int layout, OutT* __restrict__ out,
                      int64_t out_ld, int64_t out_col0, int32_t* __restrict__ rows_out,
                      int32_t* __restrict__ slots_out, int64_t slot_base, int G) {
  extern __shared__ unsigned char smem_raw[];
  int64_t* s_base = reinterpret_cast<int64_t*>(smem_raw);
  int32_t* s_vocab = reinterpret_cast<int32_t*>(s_base + K);
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    s_base[i] = row_base[i];
    s_vocab[i] = vocab[i];
  }
  __syncthreads();
  const int lane = threadIdx.x % G;
  // the lanes of this group inside its warp: groups of one warp walk bags of different lengths, so the shuffles
  // below name only the group's own lanes
  const unsigned gmask = G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int64_t n_bags = (int64_t)B * K;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  for (int64_t slot = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; slot < n_bags; slot += groups_total) {
    const int b = (int)(slot / K), k = (int)(slot - (int64_t)b * K);
    const int64_t j = layout == TT_LAYOUT_SAMPLE_MAJOR ? slot : (int64_t)k * B + b;
    const int64_t beg = __ldg(bag_offsets + j), end = __ldg(bag_offsets + j + 1);
    const int64_t hi = (int64_t)s_vocab[k] - 1, rb = s_base[k];
    for (int c0 = 0; c0 < vec_per_row; c0 += G) {  // one trip for E <= 128
      const int c = c0 + lane;
      const bool cv = c < vec_per_row;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t p0 = beg; p0 < end; p0 += G) {
        const int cnt = (int)(end - p0 < G ? end - p0 : G);
        int32_t my_row = 0;  // arena rows fit 31 bits (checked on the host)
        if (lane < cnt) {
          int64_t id = __ldg(values + p0 + lane);
          id = id < 0 ? 0 : (id > hi ? hi : id);
          my_row = (int32_t)(rb + id);
          if (c0 == 0) {
            if (rows_out) rows_out[p0 + lane] = my_row;
            if (slots_out) slots_out[p0 + lane] = (int32_t)(slot_base + slot);
          }
        }
        for (int q = 0; q < cnt; q += JAG_MLP) {
          int32_t r[JAG_MLP];
          float4 v[JAG_MLP];
#pragma unroll
          for (int u = 0; u < JAG_MLP; ++u) r[u] = __shfl_sync(gmask, my_row, (q + u) & (G - 1), G);
#pragma unroll
          for (int u = 0; u < JAG_MLP; ++u)
            if (q + u < cnt && cv) v[u] = __ldg(reinterpret_cast<const float4*>(arena + (int64_t)r[u] * E) + c);
#pragma unroll
          for (int u = 0; u < JAG_MLP; ++u)
            if (q + u < cnt && cv) f4_add(acc, v[u]);
        }
      }
      if (cv) store_vec4<OutT>(out + (int64_t)b * out_ld + out_col0 + (int64_t)k * E + 4 * c, acc);
    }
  }
}

static int lanes_per_row(int E) {
  int g = 1;
  while (g < (E >> 2) && g < 32) g <<= 1;
  return g;
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort (8-bit digits) of (row, slot) pairs by row
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
// keys per thread: 8 (2048-key tiles) for large sorts; 2 (512-key tiles) below 2^19 pairs — the reference-sized step sorts 65 k
// pairs, which is 32 long-lived CTAs at 8 keys per thread: four times as many short-lived ones finish the latency-bound chain
// sooner and do not block SMs that the persistent sim-CE kernel is waiting for
constexpr int RS_ITEMS_BIG = 8, RS_ITEMS_SMALL = 2;
static inline int rs_items_for(int64_t n) { return n < ((int64_t)1 << 19) ? RS_ITEMS_SMALL : RS_ITEMS_BIG; }

template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ d_n, int shift,
                  int32_t* __restrict__ hist, int n_tiles, int64_t* __restrict__ d_advance) {
  constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
  __shared__ int32_t s_hist[256];
  // first pass of tt_ebc_sort_rows_advance: the dedup epoch moves on inside the sort itself (no extra node in a captured step);
  // every later kernel of this sort and of the table update reads the new value in stream order
  if (d_advance && blockIdx.x == 0 && threadIdx.x == 0) *d_advance += 1;
  if (d_n) {  // device-side count (sharded tables): the launch covers the capacity, the work and the histogram pitch follow n
    n = *d_n;
    n_tiles = n > 0 ? (int)((n + RS_TILE - 1) / RS_TILE) : 1;
    if ((int)blockIdx.x >= n_tiles) return;
  }
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int64_t idx = base + (int64_t)i * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&s_hist[((uint32_t)keys[idx] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = s_hist[threadIdx.x];
}

// exclusive scan of `len` int32 counters in place, one block, the counters staged in shared memory (coalesced global traffic);
// len <= ~40k entries (the reference-sized step)
__global__ void __launch_bounds__(1024) scan_exclusive_smem_kernel(int32_t* __restrict__ data, int len, int per,
                                                                   const int32_t* __restrict__ d_n, int RS_TILE) {
  extern __shared__ int32_t sh[];
  __shared__ int32_t s_part[1024];
  const int t = threadIdx.x;
  if (d_n) {  // shorter than the launch-time bound: shared memory was sized for the bound
    len = 256 * (*d_n > 0 ? (*d_n + RS_TILE - 1) / RS_TILE : 1);
    per = (len + 1023) / 1024;
  }
  // element i lives at sh[(i / per) * (per | 1) + i % per] (odd row pitch: no bank conflicts); quotient and remainder are carried
  // incrementally — two runtime divisions per element made this single-CTA kernel the long pole of the small sorts
  const int pitch = per | 1, dq = 1024 / per, dr = 1024 % per;
  {
    int q = t / per, r = t % per;
    for (int i = t; i < len; i += 1024) {
      sh[q * pitch + r] = data[i];
      q += dq;
      r += dr;
      if (r >= per) {
        r -= per;
        ++q;
      }
    }
  }
  __syncthreads();
  int32_t* mine = sh + t * (per | 1);
  const int lo = t * per, cnt = lo >= len ? 0 : (lo + per <= len ? per : len - lo);
  int32_t sum = 0;
  for (int i = 0; i < cnt; ++i) sum += mine[i];
  s_part[t] = sum;
  __syncthreads();
  {  // inclusive scan of the 1024 thread sums: shuffles inside each warp, then the 32 warp totals by warp 0
    const int lane = t & 31, warp = t >> 5;
    int32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    __shared__ int32_t s_wtot[32];
    if (lane == 31) s_wtot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = s_wtot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_wtot[lane] = w;
    }
    __syncthreads();
    s_part[t] = x + (warp > 0 ? s_wtot[warp - 1] : 0);
  }
  int32_t run = s_part[t] - sum;
  for (int i = 0; i < cnt; ++i) {
    const int32_t c = mine[i];
    mine[i] = run;
    run += c;
  }
  __syncthreads();
  {
    int q = t / per, r = t % per;
    for (int i = t; i < len; i += 1024) {
      data[i] = sh[q * pitch + r];
      q += dq;
      r += dr;
      if (r >= per) {
        r -= per;
        ++q;
      }
    }
  }
}

// Large inputs (millions of pairs: the multi-hot configurations): the counter array no longer fits one CTA's shared memory and a
// single-CTA scan would dominate the sort. Three-kernel scan instead: per-block totals, a one-block scan of those totals, and
// a per-block exclusive scan seeded with the block's offset. SCAN_CHUNK counters per block, 128-bit loads/stores.
constexpr int SCAN_PER = 8;
constexpr int SCAN_CHUNK = 1024 * SCAN_PER;

__device__ __forceinline__ int64_t scan_len(int64_t len, const int32_t* __restrict__ d_n, int RS_TILE) {
  return d_n ? 256 * (int64_t)(*d_n > 0 ? (*d_n + RS_TILE - 1) / RS_TILE : 1) : len;
}

// block-wide inclusive scan of one value per thread (1024 threads); returns the inclusive prefix, *total = block sum
__device__ __forceinline__ int32_t block_scan_1024(int32_t x, int32_t* s_warp, int32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int32_t w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  if (total) *total = s_warp[31];
  return x + (warp > 0 ? s_warp[warp - 1] : 0);
}

__global__ void __launch_bounds__(1024)
scan_chunk_totals_kernel(const int32_t* __restrict__ data, int64_t len, const int32_t* __restrict__ d_n, int32_t* __restrict__ totals, int tile) {
  __shared__ int32_t s_warp[32];
  len = scan_len(len, d_n, tile);
  const int64_t lo = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_PER;
  int32_t sum = 0;
  if (lo + SCAN_PER <= len) {
    const int4 a = *reinterpret_cast<const int4*>(data + lo), b = *reinterpret_cast<const int4*>(data + lo + 4);
    sum = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  } else {
    for (int64_t i = lo; i < len; ++i) sum += data[i];
  }
  int32_t total;
  block_scan_1024(sum, s_warp, &total);
  if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
scan_totals_kernel(int32_t* __restrict__ totals, int64_t len, const int32_t* __restrict__ d_n, int tile) {
  __shared__ int32_t s_warp[32];
  len = scan_len(len, d_n, tile);
  const int nb = (int)((len + SCAN_CHUNK - 1) / SCAN_CHUNK);
  const int per = (nb + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
  int32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += totals[i];
  int32_t run = block_scan_1024(sum, s_warp, nullptr) - sum;
  for (int i = lo; i < hi; ++i) {
    const int32_t c = totals[i];
    totals[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(1024)
scan_chunk_apply_kernel(int32_t* __restrict__ data, int64_t len, const int32_t* __restrict__ d_n, const int32_t* __restrict__ totals, int tile) {
  __shared__ int32_t s_warp[32];
  len = scan_len(len, d_n, tile);
  const int64_t lo = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_PER;
  int32_t v[SCAN_PER];
  const bool full = lo + SCAN_PER <= len;
  if (full) {
    const int4 a = *reinterpret_cast<const int4*>(data + lo), b = *reinterpret_cast<const int4*>(data + lo + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_PER; ++i) v[i] = lo + i < len ? data[lo + i] : 0;
  }
  int32_t sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER; ++i) sum += v[i];
  int32_t run = block_scan_1024(sum, s_warp, nullptr) - sum + totals[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_PER; ++i) {
    const int32_t c = v[i];
    v[i] = run;
    run += c;
  }
  if (full) {
    *reinterpret_cast<int4*>(data + lo) = make_int4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<int4*>(data + lo + 4) = make_int4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_PER; ++i)
      if (lo + i < len) data[lo + i] = v[i];
  }
}

static int launch_scan(int32_t* data, int64_t len, const int32_t* d_n, int32_t* totals, int tile, cudaStream_t s) {
  const int per = (int)((len + 1023) / 1024);
  const size_t smem = (size_t)1024 * (per | 1) * 4;
  if (smem <= 160 * 1024) {
    TT_CUDA(cudaFuncSetAttribute(scan_exclusive_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scan_exclusive_smem_kernel<<<1, 1024, smem, s>>>(data, (int)len, per, d_n, tile);
  } else {
    const int nb = (int)((len + SCAN_CHUNK - 1) / SCAN_CHUNK);
    scan_chunk_totals_kernel<<<nb, 1024, 0, s>>>(data, len, d_n, totals, tile);
    scan_totals_kernel<<<1, 1024, 0, s>>>(totals, len, d_n, tile);
    scan_chunk_apply_kernel<<<nb, 1024, 0, s>>>(data, len, d_n, totals, tile);
  }
  return TT_OK;
}

template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int64_t n,
                     const int32_t* __restrict__ d_n, int shift, const int32_t* __restrict__ hist_scanned, int n_tiles,
                     int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
  if (d_n) {
    n = *d_n;
    n_tiles = n > 0 ? (int)((n + RS_TILE - 1) / RS_TILE) : 1;
    if ((int)blockIdx.x >= n_tiles) return;
  }
  // each warp owns 256 consecutive keys; item i of the warp is keys [i*32, i*32+32) of that chunk, so
  // (item, lane) order == index order and the ranking below is stable.
  __shared__ int32_t s_warp_hist[RS_THREADS / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
  int32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int64_t idx = base + i * 32 + lane;
    const bool ok = idx < n;
    key[i] = ok ? keys_in[idx] : 0;
    val[i] = ok ? (vals_in ? vals_in[idx] : (int32_t)idx) : 0;
    const unsigned active = __ballot_sync(0xffffffffu, ok);
    rank[i] = -1;
    if (ok) {
      const uint32_t d = ((uint32_t)key[i] >> shift) & 255u;
      const unsigned peers = __match_any_sync(active, d);
      const int leader = __ffs(peers) - 1;
      int32_t prev = 0;
      if (lane == leader) {
        prev = s_warp_hist[warp][d];
        s_warp_hist[warp][d] = prev + __popc(peers);
      }
      prev = __shfl_sync(peers, prev, leader);
      rank[i] = prev + __popc(peers & ((1u << lane) - 1u));
    }
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive scan over warps, plus this tile's global start for the digit
  {
    const int d = threadIdx.x;  // RS_THREADS == 256 digits
    int32_t run = hist_scanned[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) {
      int32_t c = s_warp_hist[w][d];
      s_warp_hist[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    if (rank[i] >= 0) {
      const uint32_t d = ((uint32_t)key[i] >> shift) & 255u;
      const int32_t pos = s_warp_hist[warp][d] + rank[i];
      keys_out[pos] = key[i];
      vals_out[pos] = val[i];
    }
  }
}

__global__ void segment_bounds_kernel(const int32_t* __restrict__ srow, int64_t n, int32_t* __restrict__ row_begin,
                                      int32_t* __restrict__ row_end, int32_t* __restrict__ row_stamp, int32_t stamp,
                                      const int64_t* __restrict__ d_stamp, const int32_t* __restrict__ d_n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d_n) n = *d_n;
  if (i >= n) return;
  if (d_stamp) stamp += (int32_t)(*d_stamp);
  const int32_t r = srow[i];
  if (i == 0 || srow[i - 1] != r) {
    row_begin[r] = (int32_t)i;
    row_stamp[r] = stamp;
  }
  if (i == n - 1 || srow[i + 1] != r) row_end[r] = (int32_t)(i + 1);
}

// ------------------------------------------------------------------------------------------------
// segment-sum + Adam
// ------------------------------------------------------------------------------------------------
constexpr int SEG_CHUNK = 64;  // rows with more than SEG_CHUNK entries are reduced in two fixed levels

struct GradSrc {
  const void* g0;
  const void* g1;
  int64_t ld0, ld1, col0, col1, n0;
  int K0, K1;
  // row-wise sharded tables: slot = src_rank * n_per_rank + local slot; the pooled-gradient buffers of every rank are
  // peer-mapped (NVLink) and listed in `table` as [rank*2 + tower] — the segment-sum pulls them directly
  const void* const* table;
  int64_t n_per_rank;
  // optional per-key column offsets (elements) inside a sample's row of the pooled-gradient buffer; NULL => col + k*E.
  // Used when only a subset of a tower's keys (the large tables) goes through the sort-dedup path.
  const int32_t* cmap0;
  const int32_t* cmap1;
};

// Branch-free on purpose (selects, one division): callers keep several of these slot -> column map -> gradient chains in flight
// per thread, and a branch between the loads of two chains makes the compiler serialise them.
template <typename GT>
__device__ __forceinline__ float4 load_grad_vec(const GradSrc& s, int32_t slot, int E, int c) {
  const void* p0 = s.g0;
  const void* p1 = s.g1;
  if (s.table) {  // kernel-uniform
    const int r = (int)(slot / s.n_per_rank);
    slot -= (int32_t)(r * s.n_per_rank);
    p0 = s.table[2 * r];
    p1 = s.table[2 * r + 1];
  }
  const bool first = slot < s.n0;
  const int sl = first ? slot : slot - (int32_t)s.n0;
  const int K = first ? s.K0 : s.K1;
  const int b = sl / K, k = sl - b * K;
  const int32_t* cm = first ? s.cmap0 : s.cmap1;
  const int64_t colk = cm ? (int64_t)__ldg(cm + k) : (first ? s.col0 : s.col1) + (int64_t)k * E;
  const GT* base = reinterpret_cast<const GT*>(first ? p0 : p1);
  const int64_t off = (int64_t)b * (first ? s.ld0 : s.ld1) + colk + 4 * c;
  if constexpr (sizeof(GT) == 4) {
    return __ldg(reinterpret_cast<const float4*>(base + off));
  } else {
    return unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(base + off)));
  }
}

// Compaction: gs[i] = the pooled-gradient vector of sorted entry i (fp32). Every thread is an independent
// index -> address -> load -> store chain, so tens of thousands of (possibly peer / NVLink) 16-byte gathers are in flight at
// once; the segment sums below then stream `gs` contiguously instead of chasing slots one dependent load at a time.
template <typename GT>
__global__ void __launch_bounds__(256)
ebc_gather_sorted_kernel(const int32_t* __restrict__ sslot, int64_t n, const int32_t* __restrict__ d_n, GradSrc src,
                         int E, float* __restrict__ gs) {
  if (d_n) n = *d_n;
  const int vec_per_row = E >> 2;
  const int64_t total = n * vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / vec_per_row;
    const int c = (int)(t - i * vec_per_row);
    *reinterpret_cast<float4*>(gs + i * E + 4 * c) = load_grad_vec<GT>(src, __ldg(sslot + i), E, c);
  }
}

// The same compaction for LOCAL sources (no peer table) and fewer than 2^31 vectors: four entries per thread, walked in phases
// (slots, column maps, gradient vectors, stores) with no branch between the loads of a phase, so that the four dependent chains of
// a thread overlap — the kernel is the first link of the table update at the tail of the train step, and latency-bound there.
template <typename GT>
__global__ void __launch_bounds__(256)
ebc_gather_sorted4_kernel(const int32_t* __restrict__ sslot, int64_t n, const int32_t* __restrict__ d_n, GradSrc src, int E,
                          float* __restrict__ gs) {
  if (d_n) n = *d_n;
  constexpr int U = 4;
  // thread -> (entry lane, 16-byte chunk c of the row); the host launches this kernel only when the chunk count divides the block
  // size, so one division serves all four entries of a thread (they are `istride` entries apart)
  const unsigned vec = (unsigned)E >> 2;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c0 = gid % vec, lane_e = gid / vec;
  const unsigned istride = gridDim.x * blockDim.x / vec;
  for (unsigned i0 = lane_e; i0 < (unsigned)n; i0 += istride * U) {
    unsigned idx[U], c[U];
    int32_t slot[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned i = i0 + u * istride;
      ok[u] = i < (unsigned)n;
      idx[u] = ok[u] ? i : 0u;
      c[u] = c0;
      slot[u] = __ldg(sslot + idx[u]);
    }
    const GT* base[U];
    int64_t off[U];
    int k[U], colk[U];
    bool first[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      first[u] = slot[u] < src.n0;
      const int sl = first[u] ? slot[u] : slot[u] - (int32_t)src.n0;
      const int K = first[u] ? src.K0 : src.K1;
      const int b = sl / K;
      k[u] = sl - b * K;
      base[u] = reinterpret_cast<const GT*>(first[u] ? src.g0 : src.g1);
      off[u] = (int64_t)b * (first[u] ? src.ld0 : src.ld1) + 4 * c[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int32_t* cm = first[u] ? src.cmap0 : src.cmap1;
      colk[u] = cm ? __ldg(cm + k[u]) : (int)(first[u] ? src.col0 : src.col1) + k[u] * E;
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const GT* p = base[u] + off[u] + colk[u];
      if constexpr (sizeof(GT) == 4) v[u] = __ldg(reinterpret_cast<const float4*>(p));
      else v[u] = unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(p)));
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (ok[u]) *reinterpret_cast<float4*>(gs + (int64_t)idx[u] * E + 4 * c[u]) = v[u];
  }
}

// Where the segment sums read the gradient vector of sorted entry i: from the compacted copy `gs` (small steps: the copy is
// L2-resident and the sums stream it), or straight from the pooled-gradient buffers through the sorted slot (`gs` NULL; large
// multi-hot steps: B*K*L compacted vectors would cost an HBM round trip while the B*K pooled gradients themselves stay in L2).
// Same values in the same order either way, so the two modes are bit-identical.
struct GradFetch {
  const float* gs;
  const int32_t* sslot;
  GradSrc src;
  int bf16;
};

template <bool DIRECT, typename GT>
__device__ __forceinline__ float4 fetch_grad_t(const GradFetch& f, int64_t i, int E, int c) {
  if (!DIRECT) return __ldg(reinterpret_cast<const float4*>(f.gs + i * E + 4 * c));
  return load_grad_vec<GT>(f.src, __ldg(f.sslot + i), E, c);
}
template <bool DIRECT>
__device__ __forceinline__ float4 fetch_grad(const GradFetch& f, int64_t i, int E, int c) {
  return (DIRECT && f.bf16) ? fetch_grad_t<DIRECT, __nv_bfloat16>(f, i, E, c) : fetch_grad_t<DIRECT, float>(f, i, E, c);
}

// acc += sum_{i in [lo,hi)} grad(i) in ascending i; batches of BATCH keep that many loads in flight before the first add. The
// ragged last batch is fetched like a full one (entries past `hi` re-read entry `lo` and are not added): a segment of up to BATCH
// entries — almost every row of a large table — costs one load latency instead of one per entry. The source dtype is dispatched
// OUTSIDE the batch: a branch between the fetches of a batch makes the compiler serialise their slot -> gradient chains.
template <bool DIRECT, typename GT, int BATCH>
__device__ __forceinline__ void gather_sum_t(float4& acc, const GradFetch& f, int64_t lo, int64_t hi, int E, int c) {
  for (int64_t i = lo; i < hi; i += BATCH) {
    float4 v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) v[u] = fetch_grad_t<DIRECT, GT>(f, i + u < hi ? i + u : lo, E, c);
#pragma unroll
    for (int u = 0; u < BATCH; ++u)
      if (i + u < hi) f4_add(acc, v[u]);
  }
}
template <bool DIRECT, int BATCH = 8>
__device__ __forceinline__ void gather_sum(float4& acc, const GradFetch& f, int64_t lo, int64_t hi, int E, int c) {
  if (DIRECT && f.bf16) gather_sum_t<DIRECT, __nv_bfloat16, BATCH>(acc, f, lo, hi, E, c);
  else gather_sum_t<DIRECT, float, BATCH>(acc, f, lo, hi, E, c);
}

// level 1 for long rows: SUB lane groups per chunk of SEG_CHUNK sorted entries. Each group sums a contiguous
// SEG_CHUNK / SUB slice of the piece in order, the SUB partial sums are folded by a fixed butterfly ((0+1)+(2+3)):
// deterministic, and SUB times the parallelism of one group per chunk (the kernel is latency bound: few chunks, long chains).
// a sub-group's whole slice (SEG_CHUNK / SUB = 16 entries at E <= 32) in one batch: the kernel is a chain of dependent round trips
// (sorted row -> row bounds -> gradients) run by few CTAs (ncu: 10 % of the warp slots), so every batch is a round trip saved
constexpr int PIECE_BATCH = 16;
template <bool DIRECT>
__global__ void __launch_bounds__(256)
ebc_piece_kernel(const int32_t* __restrict__ srow, const GradFetch gs, int64_t n,
                 const int32_t* __restrict__ d_n, const int32_t* __restrict__ row_begin,
                 const int32_t* __restrict__ row_end, int E, int G, int SUB, float* __restrict__ pieces) {
  if (d_n) n = *d_n;
  const int unit = G * SUB;  // lanes per chunk: divides 32
  const int lane = threadIdx.x % G;
  const int sub = (threadIdx.x % unit) / G;
  const int64_t chunk = (int64_t)blockIdx.x * (blockDim.x / unit) + threadIdx.x / unit;
  const int64_t lo = chunk * SEG_CHUNK;
  const bool live = lo < n;  // whole warps stay converged for the shuffles: no early return
  const int64_t hi = live ? ((lo + SEG_CHUNK < n) ? lo + SEG_CHUNK : n) : lo;
  const int vec_per_row = E >> 2;
  const int span = SEG_CHUNK / SUB;
  int64_t b0 = 0, e0 = 0, b1 = 0, e1 = 0;
  if (live) {
    const int32_t r0 = srow[lo], r1 = srow[hi - 1];
    b0 = row_begin[r0]; e0 = row_end[r0];
    b1 = row_begin[r1]; e1 = row_end[r1];
  }
  // piece 0: the long row covering index lo; piece 1: a long row that begins strictly inside the chunk (it runs past hi)
  const bool p0 = live && (e0 - b0 > SEG_CHUNK), p1 = live && (e1 - b1 > SEG_CHUNK) && b1 > lo;
  const int64_t pe = e0 < hi ? e0 : hi;
  for (int c0 = 0; c0 < vec_per_row; c0 += G) {  // uniform trip count: every lane of the warp reaches the shuffles
    const int c = c0 + lane;
    const bool cv = c < vec_per_row;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 && cv) {
      const int64_t s_lo = lo + (int64_t)sub * span, s_hi = s_lo + span < pe ? s_lo + span : pe;
      if (s_lo < s_hi) gather_sum<DIRECT, PIECE_BATCH>(a0, gs, s_lo, s_hi, E, c);
    }
    if (p1 && cv) {
      const int64_t s_lo = b1 + (int64_t)sub * span, s_hi = s_lo + span < hi ? s_lo + span : hi;
      if (s_lo < s_hi) gather_sum<DIRECT, PIECE_BATCH>(a1, gs, s_lo, s_hi, E, c);
    }
    for (int off = G; off < unit; off <<= 1) {
      a0.x += __shfl_xor_sync(0xffffffffu, a0.x, off); a0.y += __shfl_xor_sync(0xffffffffu, a0.y, off);
      a0.z += __shfl_xor_sync(0xffffffffu, a0.z, off); a0.w += __shfl_xor_sync(0xffffffffu, a0.w, off);
      a1.x += __shfl_xor_sync(0xffffffffu, a1.x, off); a1.y += __shfl_xor_sync(0xffffffffu, a1.y, off);
      a1.z += __shfl_xor_sync(0xffffffffu, a1.z, off); a1.w += __shfl_xor_sync(0xffffffffu, a1.w, off);
    }
    if (sub == 0 && cv) {
      if (p0) *reinterpret_cast<float4*>(pieces + ((chunk * 2 + 0) * E) + 4 * c) = a0;
      if (p1) *reinterpret_cast<float4*>(pieces + ((chunk * 2 + 1) * E) + 4 * c) = a1;
    }
  }
}

template <bool DIRECT>
__device__ __forceinline__ float4 row_grad_sum(const GradFetch& gs, int64_t b, int64_t e, int E, int c,
                                               const float* __restrict__ pieces) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e - b <= SEG_CHUNK) {
    gather_sum<DIRECT>(acc, gs, b, e, E, c);
  } else {
    for (int64_t ch = b / SEG_CHUNK; ch <= (e - 1) / SEG_CHUNK; ++ch) {
      const int which = (b <= ch * SEG_CHUNK) ? 0 : 1;
      f4_add(acc, *reinterpret_cast<const float4*>(pieces + ((ch * 2 + which) * E) + 4 * c));
    }
  }
  return acc;
}

// one lane group per arena row, every row (dense-exact Adam and/or dense gradient materialisation)
template <bool DO_ADAM, bool WRITE_GRAD, bool DIRECT>
__global__ void __launch_bounds__(256, 3)
ebc_rows_dense_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, float* __restrict__ gout,
                      int64_t total_rows, int E, int G, const GradFetch gs,
                      const int32_t* __restrict__ row_begin, const int32_t* __restrict__ row_end,
                      const int32_t* __restrict__ row_stamp, int32_t stamp, const int64_t* __restrict__ d_stamp,
                      const float* __restrict__ pieces, const tt_adam_hyper* __restrict__ d_hyper, int untouched_only) {
  if (d_stamp) stamp += (int32_t)(*d_stamp);
  tt_adam_hyper h;
  if (DO_ADAM) h = *d_hyper;
  const int lane = threadIdx.x % G;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; r < total_rows; r += groups_total) {
    const bool hit = row_stamp[r] == stamp;
    if (untouched_only && hit) continue;  // the touched rows are updated later, once their gradients exist
    int64_t b = 0, e = 0;
    if (hit) {
      b = row_begin[r];
      e = row_end[r];
    }
    for (int c = lane; c < vec_per_row; c += G) {
      const int64_t off = r * E + 4 * c;
      // issue the streaming w/m/v loads first: they overlap with the dependent stamp -> bounds -> gradient chain
      float4 wv, mv, vv;
      if (DO_ADAM) {
        wv = *reinterpret_cast<const float4*>(w + off);
        mv = *reinterpret_cast<const float4*>(m + off);
        vv = *reinterpret_cast<const float4*>(v + off);
      }
      float4 g = hit ? row_grad_sum<DIRECT>(gs, b, e, E, c, pieces) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (WRITE_GRAD) *reinterpret_cast<float4*>(gout + off) = g;
      if (DO_ADAM) {
        adam_vec4(wv, mv, vv, g, h);
        *reinterpret_cast<float4*>(w + off) = wv;
        *reinterpret_cast<float4*>(m + off) = mv;
        *reinterpret_cast<float4*>(v + off) = vv;
      }
    }
  }
}

// dense-exact Adam over an arena that is much larger than the L2 (the 10 M-row / 500 M-row configurations): pure streaming of
// w/m/v (24 B per element). Two rows per lane group and iteration keep six 128-bit loads per lane in flight ahead of the
// stamp -> bounds -> gradient chain; loads and stores are evict-first so the sorted gradients survive in the L2.
template <bool DIRECT>
__global__ void __launch_bounds__(256, 3)
ebc_rows_dense_stream_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, int64_t total_rows, int E, int G,
                             const GradFetch gs, const int32_t* __restrict__ row_begin, const int32_t* __restrict__ row_end,
                             const int32_t* __restrict__ row_stamp, int32_t stamp, const int64_t* __restrict__ d_stamp,
                             const float* __restrict__ pieces, const tt_adam_hyper* __restrict__ d_hyper) {
  if (d_stamp) stamp += (int32_t)(*d_stamp);
  const tt_adam_hyper h = *d_hyper;
  const int lane = threadIdx.x % G;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  for (int64_t r0 = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; r0 < total_rows; r0 += 2 * groups_total) {
    const int64_t r1 = r0 + groups_total;
    const bool two = r1 < total_rows;
    for (int c = lane; c < vec_per_row; c += G) {
      const int64_t o0 = r0 * E + 4 * c, o1 = r1 * E + 4 * c;
      float4 w0 = __ldcs(reinterpret_cast<const float4*>(w + o0)), m0 = __ldcs(reinterpret_cast<const float4*>(m + o0)),
             v0 = __ldcs(reinterpret_cast<const float4*>(v + o0));
      float4 w1, m1, v1;
      if (two) {
        w1 = __ldcs(reinterpret_cast<const float4*>(w + o1));
        m1 = __ldcs(reinterpret_cast<const float4*>(m + o1));
        v1 = __ldcs(reinterpret_cast<const float4*>(v + o1));
      }
      const bool hit0 = __ldg(row_stamp + r0) == stamp, hit1 = two && __ldg(row_stamp + r1) == stamp;
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      if (hit0) g0 = row_grad_sum<DIRECT>(gs, row_begin[r0], row_end[r0], E, c, pieces);
      if (hit1) g1 = row_grad_sum<DIRECT>(gs, row_begin[r1], row_end[r1], E, c, pieces);
      adam_vec4(w0, m0, v0, g0, h);
      __stcs(reinterpret_cast<float4*>(w + o0), w0);
      __stcs(reinterpret_cast<float4*>(m + o0), m0);
      __stcs(reinterpret_cast<float4*>(v + o0), v0);
      if (two) {
        adam_vec4(w1, m1, v1, g1, h);
        __stcs(reinterpret_cast<float4*>(w + o1), w1);
        __stcs(reinterpret_cast<float4*>(m + o1), m1);
        __stcs(reinterpret_cast<float4*>(v + o1), v1);
      }
    }
  }
}

// Bring every row up to the current step (before the parameters are read from outside: state_dict / checkpoints, and before the
// history ring would wrap). One lane group per row; rows that are current cost one 4-byte read.
__global__ void __launch_bounds__(256)
ebc_lazy_flush_kernel(const tt_lazy_ctx* __restrict__ lazy, int64_t rows, int E, int G) {
  const tt_lazy_ctx L = *lazy;
  const tt_adam_hyper h = *L.hyper;
  const int lane = threadIdx.x % G;
  const unsigned gmask = G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  float* w = const_cast<float*>(L.arena_w);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; r < rows; r += groups_total) {
    const long long last = L.row_last[r];
    if (last >= h.step) continue;
    for (int c = lane; c < vec_per_row; c += G) {
      const int64_t off = r * E + 4 * c;
      float4 wv = *reinterpret_cast<const float4*>(w + off), mv = *reinterpret_cast<const float4*>(L.arena_m + off),
             vv = *reinterpret_cast<const float4*>(L.arena_v + off);
      lazy_replay_vec4(wv, mv, vv, last + 1, h.step, L, h);
      *reinterpret_cast<float4*>(w + off) = wv;
      *reinterpret_cast<float4*>(L.arena_m + off) = mv;
      *reinterpret_cast<float4*>(L.arena_v + off) = vv;
    }
    __syncwarp(gmask);
    if (lane == 0) L.row_last[r] = (int32_t)h.step;
  }
}

// sparse mode: lane groups walk the sorted entries; only segment heads work. HBM-bound at scale (millions of touched rows, each
// a random 3 x E*4-byte read-modify-write), so the kernel is built around memory-level parallelism:
//   * persistent grid-stride CTAs, SP_ROWS entries per lane group and iteration. The sorted rows of the NEXT iteration (the entry,
//     its predecessor for the head test, its successor for the segment end) are fetched while this iteration's rows are in
//     flight, so the dependent chain of an iteration is ONE memory round trip: w / m / v / gradient of every entry leave together,
//     with no branch between them (an entry that is not a head reads row 0 and stores nothing);
//   * a head finds its segment end from that successor (most segments are one entry; longer ones scan forward through the
//     sorted rows: sequential, cached) instead of a random row_end[r] lookup;
//   * w/m/v accesses are streaming (evict-first), so the pooled/compacted gradients keep the L2.
constexpr int SP_ROWS = 2;
template <bool DIRECT, bool LAZY>
__global__ void __launch_bounds__(256, 3)
ebc_rows_sparse_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, int E, int G,
                       const int32_t* __restrict__ srow, const GradFetch gs, int64_t n,
                       const int32_t* __restrict__ d_n, const int32_t* __restrict__ row_end,
                       const float* __restrict__ pieces, const tt_adam_hyper* __restrict__ d_hyper,
                       const tt_lazy_ctx* __restrict__ lazy) {
  if (d_n) n = *d_n;
  const tt_adam_hyper h = *d_hyper;
  tt_lazy_ctx L = {};
  if (LAZY) L = *lazy;
  const unsigned gmask = G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
  const int lane = threadIdx.x % G;
  const int vec_per_row = E >> 2;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int32_t n32 = (int32_t)n;  // sorted positions fit 31 bits (tt_ebc_sort_rows checks n)
  // (row, predecessor's row, successor's row) of the SP_ROWS entries of an iteration; -1 / -2 / -3 where there is none
  auto fetch_rows = [&](int64_t base, int32_t (&rr)[SP_ROWS], int32_t (&rp)[SP_ROWS], int32_t (&rn)[SP_ROWS]) {
#pragma unroll
    for (int u = 0; u < SP_ROWS; ++u) {
      const int64_t iu = base + u * groups_total;
      const bool in = iu < n;
      const int32_t i = in ? (int32_t)iu : 0;
      const int32_t a = __ldg(srow + i), p = __ldg(srow + (i > 0 ? i - 1 : 0)), x = __ldg(srow + (i + 1 < n32 ? i + 1 : i));
      rr[u] = in ? a : -1;
      rp[u] = (in && i > 0) ? p : -2;
      rn[u] = (in && i + 1 < n32) ? x : -3;
    }
  };
  int32_t r_nx[SP_ROWS], p_nx[SP_ROWS], x_nx[SP_ROWS];
  int64_t base = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
  if (base < n) fetch_rows(base, r_nx, p_nx, x_nx);
  for (; base < n; base += SP_ROWS * groups_total) {
    int32_t i[SP_ROWS], e[SP_ROWS], r[SP_ROWS], nx[SP_ROWS];
    bool head[SP_ROWS];
    int32_t last[SP_ROWS];
    float4 wv[SP_ROWS], mv[SP_ROWS], vv[SP_ROWS], g0[SP_ROWS];
#pragma unroll
    for (int u = 0; u < SP_ROWS; ++u) {
      i[u] = (int32_t)(base + u * groups_total);
      head[u] = r_nx[u] >= 0 && p_nx[u] != r_nx[u];
      r[u] = head[u] ? r_nx[u] : 0;
      nx[u] = x_nx[u];
    }
    // one round trip: state and the head's own gradient entry of every row of this iteration (first — for E <= 128 only —
    // vector of the row), branch-free ...
    const int c0 = lane < vec_per_row ? lane : 0;
#pragma unroll
    for (int u = 0; u < SP_ROWS; ++u) {
      const int64_t off = (int64_t)r[u] * E + 4 * c0;
      wv[u] = __ldcs(reinterpret_cast<const float4*>(w + off));
      mv[u] = __ldcs(reinterpret_cast<const float4*>(m + off));
      vv[u] = __ldcs(reinterpret_cast<const float4*>(v + off));
      g0[u] = fetch_grad<DIRECT>(gs, head[u] ? i[u] : 0, E, c0);
      last[u] = 0;
      if (LAZY) last[u] = L.row_last[r[u]];
    }
    // ... and, underneath it, the sorted rows of the next iteration
    const int64_t base_nx = base + SP_ROWS * groups_total;
    if (base_nx < n) fetch_rows(base_nx, r_nx, p_nx, x_nx);
#pragma unroll
    for (int u = 0; u < SP_ROWS; ++u) {
      e[u] = i[u] + 1;
      if (head[u] && nx[u] == r[u]) {  // a segment of several entries: scan on (sequential reads of the sorted rows)
        int32_t x = i[u] + 2;
        while (x < n32 && x - i[u] <= SEG_CHUNK && __ldg(srow + x) == r[u]) ++x;
        e[u] = (x - i[u] > SEG_CHUNK) ? row_end[r[u]] : x;  // long rows were reduced to pieces: take the recorded end
      }
    }
    for (int c = lane; c < vec_per_row; c += G) {
#pragma unroll
      for (int u = 0; u < SP_ROWS; ++u)
        if (head[u]) {
          const int64_t off = (int64_t)r[u] * E + 4 * c;
          if (c != lane) {
            wv[u] = __ldcs(reinterpret_cast<const float4*>(w + off));
            mv[u] = __ldcs(reinterpret_cast<const float4*>(m + off));
            vv[u] = __ldcs(reinterpret_cast<const float4*>(v + off));
            g0[u] = fetch_grad<DIRECT>(gs, i[u], E, c);
          }
          float4 g;
          if (e[u] - i[u] <= SEG_CHUNK) {  // 0 + g0 + ... == the plain ascending sum, bit for bit
            g = g0[u];
            if (e[u] - i[u] > 1) gather_sum<DIRECT, DIRECT ? 2 : 8>(g, gs, (int64_t)i[u] + 1, e[u], E, c);  // compacted: a contiguous stream
          } else {  // long row: its per-chunk pieces, in chunk order (same as row_grad_sum)
            g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int32_t ch = i[u] / SEG_CHUNK; ch <= (e[u] - 1) / SEG_CHUNK; ++ch)
              f4_add(g, *reinterpret_cast<const float4*>(pieces + (((int64_t)ch * 2 + (i[u] <= ch * SEG_CHUNK ? 0 : 1)) * E) + 4 * c));
          }
          if (LAZY) lazy_replay_vec4(wv[u], mv[u], vv[u], (long long)last[u] + 1, h.step - 1, L, h);  // the steps this row sat out
          adam_vec4(wv[u], mv[u], vv[u], g, h);
          __stcs(reinterpret_cast<float4*>(w + off), wv[u]);
          __stcs(reinterpret_cast<float4*>(m + off), mv[u]);
          __stcs(reinterpret_cast<float4*>(v + off), vv[u]);
        }
    }
    if (LAZY) {  // the lanes of a group took the same path (group-uniform control flow) and have all read row_last by now
      __syncwarp(gmask);
#pragma unroll
      for (int u = 0; u < SP_ROWS; ++u)
        if (head[u] && lane == 0) L.row_last[r[u]] = (int32_t)h.step;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small tables: dense one-hot reduction instead of sort-dedup
// ------------------------------------------------------------------------------------------------
// Tables with a handful of rows (24 of the 32 notice keys have <= 25) receive every sample of the batch: hundreds to
// thousands of gradient vectors per row. Sorting them is wasted work; instead each CTA reduces a fixed chunk of samples
// into per-warp shared-memory accumulators [V_k][E] (fixed sample order per warp, warps combined in fixed order) and a
// second pass sums the chunk partials in order: deterministic, fully coalesced, no atomics.
struct SmallTab {
  int key;        // key index inside the tower
  int vocab;      // rows
  int col;        // column offset (elements) of the key's slice in the pooled-gradient row
  int64_t g_off;  // offset (floats) of the table's gradient inside gout
};
constexpr int SM_CHUNK = 128;  // smallest number of samples per CTA (sizes the partials workspace)

template <typename GT>
__global__ void __launch_bounds__(256)
ebc_small_partial_kernel(const int64_t* __restrict__ values, int B, int K, int layout, const GT* __restrict__ grad, int64_t ld,
                         int E, const SmallTab* __restrict__ tabs, float* __restrict__ part, int64_t part_stride, int chunk) {
  extern __shared__ float s_acc[];  // [8 warps][vocab * E]
  const SmallTab tb = tabs[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ve = tb.vocab * E;
  for (int i = threadIdx.x; i < 8 * ve; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  float* mine = s_acc + warp * ve;
  const int b_lo = blockIdx.y * chunk;
  const int b_hi = min(b_lo + chunk, B);
  if (E <= 32) {
    // eight samples of this warp per round: their ids (one lane each) and gradient vectors (one element per lane) are all in
    // flight together; the shared-memory accumulation then runs in ascending sample order (same order as the plain loop)
    for (int base = b_lo + warp; base < b_hi; base += 64) {
      int my_id = 0;
      {
        const int b = base + lane * 8;
        if (lane < 8 && b < b_hi) {
          int64_t id = __ldg(values + (layout == TT_LAYOUT_SAMPLE_MAJOR ? (int64_t)b * K + tb.key : (int64_t)tb.key * B + b));
          my_id = (int)(id < 0 ? 0 : (id > tb.vocab - 1 ? tb.vocab - 1 : id));
        }
      }
      float gv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = base + u * 8;
        gv[u] = (b < b_hi && lane < E) ? ldf<GT>(grad + (int64_t)b * ld + tb.col + lane) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int id = __shfl_sync(0xffffffffu, my_id, u);
        if (base + u * 8 < b_hi && lane < E) mine[id * E + lane] += gv[u];
      }
    }
  } else {
    for (int b = b_lo + warp; b < b_hi; b += 8) {
      int64_t id = __ldg(values + (layout == TT_LAYOUT_SAMPLE_MAJOR ? (int64_t)b * K + tb.key : (int64_t)tb.key * B + b));
      id = id < 0 ? 0 : (id > tb.vocab - 1 ? tb.vocab - 1 : id);
      const GT* g = grad + (int64_t)b * ld + tb.col;
      for (int e = lane; e < E; e += 32) mine[id * E + e] += ldf<GT>(g + e);
    }
  }
  __syncthreads();
  float* out = part + (int64_t)blockIdx.y * part_stride + tb.g_off;
  for (int i = threadIdx.x; i < ve; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_acc[w * ve + i];
    out[i] = t;
  }
}

__global__ void ebc_small_reduce_kernel(const float* __restrict__ part, int64_t part_stride, int chunks, int64_t n,
                                        float* __restrict__ gout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  int c = 0;
  for (; c + 8 <= chunks; c += 8) {  // eight chunk partials in flight, added in chunk order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(c + u) * part_stride + i];
#pragma unroll
    for (int u = 0; u < 8; ++u) t += v[u];
  }
  for (; c < chunks; ++c) t += part[(int64_t)c * part_stride + i];
  gout[i] = t;
}

__global__ void __launch_bounds__(256)
dense_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, const tt_adam_hyper* __restrict__ d_hyper) {
  const tt_adam_hyper h = *d_hyper;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 wv = reinterpret_cast<float4*>(w)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    adam_vec4(wv, mv, vv, gv, h);
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_elem(w[i], m[i], v[i], g[i], h);
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace tt

using namespace tt;

extern "C" {

static int ebc_forward_impl(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values,
                            const int64_t* bag_offsets, int64_t n_values, int32_t B, int32_t K, int32_t E, int32_t layout,
                            void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0, int32_t* rows_out,
                            int32_t* slots_out, int64_t slot_base, const int32_t* large_rank, int32_t K_large, int64_t row_sub,
                            tt_stream_t stream, const tt_lazy_ctx* d_lazy = nullptr, int upto_off = 0) {
  TT_CHECK_ARG(arena && d_row_base && d_vocab && values && out, "tt_ebc_forward: NULL pointer");
  TT_CHECK_ARG(!d_lazy || (bag_offsets == nullptr && large_rank), "tt_ebc_forward: lazy catch-up is implemented for the one-id-per-bag mixed lookup");
  TT_CHECK_ARG(B >= 0 && K > 0 && E > 0 && (E % 4) == 0, "tt_ebc_forward: need K>0, E%%4==0 (B=%d K=%d E=%d)", B, K, E);
  TT_CHECK_ARG(layout == TT_LAYOUT_SAMPLE_MAJOR || layout == TT_LAYOUT_KEY_MAJOR, "tt_ebc_forward: bad layout %d", layout);
  TT_CHECK_ARG(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_ebc_forward: bad out_dtype %d", out_dtype);
  TT_CHECK_ARG((out_ld % 4) == 0 && (out_col0 % 4) == 0 && out_ld >= out_col0 + (int64_t)K * E,
               "tt_ebc_forward: out_ld/out_col0 must be multiples of 4 and fit K*E columns");
  TT_CHECK_ARG(bag_offsets != nullptr || n_values == (int64_t)B * K,
               "tt_ebc_forward: without bag_offsets n_values (%lld) must equal B*K (%lld)", (long long)n_values,
               (long long)B * K);
  if (B == 0) return TT_OK;
  TT_CHECK_ARG((int64_t)B * K < ((int64_t)1 << 31), "tt_ebc_forward: B*K must be below 2^31");
  const int G = lanes_per_row(E);
  const int threads = 256;
  const int64_t n_bags = (int64_t)B * K;
  const size_t smem = (size_t)K * (sizeof(int64_t) + 2 * sizeof(int32_t));
  cudaStream_t s = (cudaStream_t)stream;
  const int groups_per_block = threads / G;
  if (bag_offsets == nullptr) {
    constexpr int UNROLL = 4;
    int64_t want = ceil_div(n_bags, (int64_t)groups_per_block * UNROLL);
    // (capping the grid at 5 CTAs per SM to leave slots for the concurrent dense-projection GEMMs was measured: the lookup went
    // from 18.8 to 31.9 us and the step got slower)
    const int64_t gcap = (int64_t)sm_count() * 16;
    int grid = (int)(want < 1 ? 1 : (want > gcap ? gcap : want));
#define TT_L1_LAUNCH(OUT_T, LZ)                                                                                                  \
  ebc_fwd_l1_kernel<OUT_T, UNROLL, LZ><<<grid, threads, smem, s>>>(arena, d_row_base, d_vocab, values, B, K, E, layout, (OUT_T*)out, \
                                                                   out_ld, out_col0, rows_out, slots_out, slot_base, G, large_rank, \
                                                                   K_large, row_sub, d_lazy, upto_off)
    if (out_dtype == TT_F32) {
      if (d_lazy) TT_L1_LAUNCH(float, true); else TT_L1_LAUNCH(float, false);
    } else {
      if (d_lazy) TT_L1_LAUNCH(__nv_bfloat16, true); else TT_L1_LAUNCH(__nv_bfloat16, false);
    }
#undef TT_L1_LAUNCH
  } else {
    int64_t want = ceil_div(n_bags, groups_per_block);
    int grid = (int)(want < 1 ? 1 : (want > (int64_t)sm_count() * 16 ? (int64_t)sm_count() * 16 : want));
    if (out_dtype == TT_F32)
      ebc_fwd_jagged_kernel<float><<<grid, threads, smem, s>>>(arena, d_row_base, d_vocab, values, bag_offsets, B, K, E,
                                                               layout, (float*)out, out_ld, out_col0, rows_out,
                                                               slots_out, slot_base, G);
    else
      ebc_fwd_jagged_kernel<__nv_bfloat16><<<grid, threads, smem, s>>>(arena, d_row_base, d_vocab, values, bag_offsets,
                                                                       B, K, E, layout, (__nv_bfloat16*)out, out_ld,
                                                                       out_col0, rows_out, slots_out, slot_base, G);
  }
  TT_LAUNCH_CHECK("ebc_fwd_kernel");
  return TT_OK;
}

int tt_ebc_forward(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values,
                   const int64_t* bag_offsets, int64_t n_values, int32_t B, int32_t K, int32_t E, int32_t layout,
                   void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0, int32_t* rows_out,
                   int32_t* slots_out, int64_t slot_base, tt_stream_t stream) {
  return ebc_forward_impl(arena, d_row_base, d_vocab, values, bag_offsets, n_values, B, K, E, layout, out, out_dtype, out_ld,
                          out_col0, rows_out, slots_out, slot_base, nullptr, 0, 0, stream);
}

int tt_ebc_forward_mixed(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values, int32_t B,
                         int32_t K, int32_t E, int32_t layout, void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0,
                         const int32_t* d_large_rank, int32_t K_large, int64_t row_sub, int32_t* rows_out, tt_stream_t stream) {
  TT_CHECK_ARG(d_large_rank && K_large >= 0 && K_large <= K, "tt_ebc_forward_mixed: bad large-table map");
  return ebc_forward_impl(arena, d_row_base, d_vocab, values, nullptr, (int64_t)B * K, B, K, E, layout, out, out_dtype, out_ld,
                          out_col0, rows_out, nullptr, 0, d_large_rank, K_large, row_sub, stream);
}

int tt_ebc_forward_mixed_lazy(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values, int32_t B,
                              int32_t K, int32_t E, int32_t layout, void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0,
                              const int32_t* d_large_rank, int32_t K_large, int64_t row_sub, int32_t* rows_out,
                              const tt_lazy_ctx* d_lazy, int32_t upto_offset, tt_stream_t stream) {
  TT_CHECK_ARG(d_large_rank && K_large >= 0 && K_large <= K && d_lazy && (upto_offset == 0 || upto_offset == -1),
               "tt_ebc_forward_mixed_lazy: bad arguments");
  return ebc_forward_impl(arena, d_row_base, d_vocab, values, nullptr, (int64_t)B * K, B, K, E, layout, out, out_dtype, out_ld,
                          out_col0, rows_out, nullptr, 0, d_large_rank, K_large, row_sub, stream, d_lazy, upto_offset);
}

int64_t tt_ebc_sort_workspace_bytes(int64_t n, int64_t total_rows) {
  (void)total_rows;
  if (n < 0) return 0;
  const int64_t tile = RS_THREADS * rs_items_for(n);
  const int64_t tiles = (n + tile - 1) / tile + 1;
  return 2 * align_up(n * 4 + 16, 256) + align_up(256 * tiles * 4, 256) + align_up((256 * tiles / SCAN_CHUNK + 2) * 4, 256);
}

static int sort_rows_impl(const int32_t* rows, const int32_t* slots, int64_t n, int64_t total_rows, void* workspace,
                          int64_t workspace_bytes, int32_t* sorted_rows, int32_t* sorted_slots, int32_t* row_begin,
                          int32_t* row_end, int32_t* row_stamp, int32_t stamp, const int64_t* d_stamp,
                          const int32_t* d_n, tt_stream_t stream, int64_t* d_advance) {
  TT_CHECK_ARG(rows && sorted_rows && sorted_slots && row_begin && row_end && row_stamp, "tt_ebc_sort_rows: NULL pointer");
  TT_CHECK_ARG(n >= 0 && n < (1ll << 31) && total_rows > 0 && total_rows < (1ll << 31), "tt_ebc_sort_rows: bad sizes");
  TT_CHECK_ARG(workspace_bytes >= tt_ebc_sort_workspace_bytes(n, total_rows) && (workspace || n == 0),
               "tt_ebc_sort_rows: workspace too small (%lld < %lld)", (long long)workspace_bytes,
               (long long)tt_ebc_sort_workspace_bytes(n, total_rows));
  if (n == 0) return TT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int bits = 0;
  while (((int64_t)1 << bits) < total_rows) ++bits;
  int passes = (bits + 7) / 8;
  if (passes < 1) passes = 1;
  const int items = rs_items_for(n), tile = RS_THREADS * items;
  const int n_tiles = (int)((n + tile - 1) / tile);
  char* ws = (char*)workspace;
  int32_t* tmp_keys = (int32_t*)ws;
  int32_t* tmp_vals = (int32_t*)(ws + align_up(n * 4 + 16, 256));
  int32_t* hist = (int32_t*)(ws + 2 * align_up(n * 4 + 16, 256));
  int32_t* scan_totals = (int32_t*)((char*)hist + align_up(256 * (int64_t)((n + tile - 1) / tile + 1) * 4, 256));
  const int32_t* kin = rows;
  const int32_t* vin = slots;  // NULL => identity on the first pass
  for (int p = 0; p < passes; ++p) {
    const bool to_final = ((passes - 1 - p) % 2) == 0;
    int32_t* kout = to_final ? sorted_rows : tmp_keys;
    int32_t* vout = to_final ? sorted_slots : tmp_vals;
    int64_t* adv = p == 0 ? d_advance : nullptr;
    if (items == RS_ITEMS_SMALL) radix_hist_kernel<RS_ITEMS_SMALL><<<n_tiles, RS_THREADS, 0, s>>>(kin, n, d_n, p * 8, hist, n_tiles, adv);
    else radix_hist_kernel<RS_ITEMS_BIG><<<n_tiles, RS_THREADS, 0, s>>>(kin, n, d_n, p * 8, hist, n_tiles, adv);
    { int rc = launch_scan(hist, (int64_t)256 * n_tiles, d_n, scan_totals, tile, s); if (rc != TT_OK) return rc; }
    if (items == RS_ITEMS_SMALL)
      radix_scatter_kernel<RS_ITEMS_SMALL><<<n_tiles, RS_THREADS, 0, s>>>(kin, vin, n, d_n, p * 8, hist, n_tiles, kout, vout);
    else
      radix_scatter_kernel<RS_ITEMS_BIG><<<n_tiles, RS_THREADS, 0, s>>>(kin, vin, n, d_n, p * 8, hist, n_tiles, kout, vout);
    kin = kout;
    vin = vout;
  }
  segment_bounds_kernel<<<ceil_div(n, 256), 256, 0, s>>>(sorted_rows, n, row_begin, row_end, row_stamp, stamp, d_stamp, d_n);
  TT_LAUNCH_CHECK("ebc_sort_rows");
  return TT_OK;
}

int tt_ebc_sort_rows(const int32_t* rows, const int32_t* slots, int64_t n, int64_t total_rows, void* workspace,
                     int64_t workspace_bytes, int32_t* sorted_rows, int32_t* sorted_slots, int32_t* row_begin,
                     int32_t* row_end, int32_t* row_stamp, int32_t stamp, const int64_t* d_stamp,
                     const int32_t* d_n, tt_stream_t stream) {
  return sort_rows_impl(rows, slots, n, total_rows, workspace, workspace_bytes, sorted_rows, sorted_slots, row_begin, row_end,
                        row_stamp, stamp, d_stamp, d_n, stream, nullptr);
}

int tt_ebc_sort_rows_advance(const int32_t* rows, const int32_t* slots, int64_t n, int64_t total_rows, void* workspace,
                             int64_t workspace_bytes, int32_t* sorted_rows, int32_t* sorted_slots, int32_t* row_begin,
                             int32_t* row_end, int32_t* row_stamp, int64_t* d_stamp, const int32_t* d_n,
                             tt_stream_t stream) {
  TT_CHECK_ARG(d_stamp && n > 0, "tt_ebc_sort_rows_advance: needs a device epoch counter and n > 0");
  return sort_rows_impl(rows, slots, n, total_rows, workspace, workspace_bytes, sorted_rows, sorted_slots, row_begin, row_end,
                        row_stamp, 0, d_stamp, d_n, stream, d_stamp);
}

int64_t tt_ebc_piece_workspace_bytes(int64_t n, int32_t E) {
  // [pieces: 2 per SEG_CHUNK of sorted entries][gs: the compacted pooled gradients, one fp32 vector per sorted entry]
  const int64_t chunks = (n + SEG_CHUNK - 1) / SEG_CHUNK + 1;
  return align_up(chunks * 2 * (int64_t)E * 4, 256) + (n > 0 ? n : 0) * (int64_t)E * 4;
}

static int backward_adam_impl(float* arena_w, float* arena_m, float* arena_v, float* grad_arena, int64_t total_rows,
                         int32_t E, const int32_t* sorted_rows, const int32_t* sorted_slots, int64_t n,
                         const int32_t* row_begin, const int32_t* row_end, const int32_t* row_stamp, int32_t stamp,
                         const int64_t* d_stamp, const void* grad0, int64_t ld0, int64_t col0_0, int32_t K0, int64_t n0, const void* grad1,
                         int64_t ld1, int64_t col0_1, int32_t K1, int64_t n1, int32_t grad_dtype, void* piece_ws,
                         int64_t piece_ws_bytes, const tt_adam_hyper* d_hyper, int32_t mode, const int32_t* d_n,
                         const void* const* grad_table, int64_t n_per_rank, const int32_t* d_colmap0, const int32_t* d_colmap1,
                         tt_stream_t stream, const tt_lazy_ctx* d_lazy) {
  const bool do_adam = arena_w != nullptr;
  TT_CHECK_ARG(do_adam || grad_arena, "tt_ebc_backward_adam: nothing to do (arena_w and grad_arena both NULL)");
  TT_CHECK_ARG(!do_adam || (arena_m && arena_v && d_hyper), "tt_ebc_backward_adam: Adam needs m, v and d_hyper");
  TT_CHECK_ARG(E > 0 && (E % 4) == 0 && total_rows > 0, "tt_ebc_backward_adam: bad E/total_rows");
  TT_CHECK_ARG(n == 0 || (sorted_rows && sorted_slots && row_begin && row_end && row_stamp && (grad0 || grad_table)),
               "tt_ebc_backward_adam: NULL index/grad pointer");
  TT_CHECK_ARG(n0 + n1 >= 0 && (n1 == 0 || grad1 || grad_table), "tt_ebc_backward_adam: grad1 missing");
  TT_CHECK_ARG(!grad_table || n_per_rank == n0 + n1, "tt_ebc_backward_adam: n_per_rank must equal n0 + n1 with a grad table");
  TT_CHECK_ARG(K0 > 0 && (n1 == 0 || K1 > 0), "tt_ebc_backward_adam: bad K");
  TT_CHECK_ARG((ld0 % 4) == 0 && (col0_0 % 4) == 0 && (ld1 % 4) == 0 && (col0_1 % 4) == 0,
               "tt_ebc_backward_adam: grad ld/col0 must be multiples of 4");
  TT_CHECK_ARG(grad_dtype == TT_F32 || grad_dtype == TT_BF16, "tt_ebc_backward_adam: bad grad dtype");
  TT_CHECK_ARG(mode == TT_ADAM_DENSE_EXACT || mode == TT_ADAM_SPARSE || mode == TT_ADAM_DENSE_UNTOUCHED, "tt_ebc_backward_adam: bad mode %d", mode);
  if (mode == TT_ADAM_DENSE_UNTOUCHED) {
    // first half of a split exact-dense step: every row the batch does NOT touch (g = wd*w), needs only the stamps of the sort
    TT_CHECK_ARG(do_adam && !grad_arena && row_stamp, "tt_ebc_backward_adam: TT_ADAM_DENSE_UNTOUCHED needs w/m/v, d_hyper and row_stamp");
    const int G0 = lanes_per_row(E);
    int64_t want0 = ceil_div(total_rows, 256 / G0);
    const int64_t cap0 = (int64_t)sm_count() * 32;
    const GradFetch none{nullptr, nullptr, GradSrc{}, 0};
    ebc_rows_dense_kernel<true, false, false><<<(int)(want0 > cap0 ? cap0 : want0), 256, 0, (cudaStream_t)stream>>>(
        arena_w, arena_m, arena_v, nullptr, total_rows, E, G0, none, row_begin, row_end, row_stamp, stamp, d_stamp, nullptr, d_hyper, 1);
    TT_LAUNCH_CHECK("ebc_rows_dense_kernel(untouched)");
    return TT_OK;
  }
  TT_CHECK_ARG(!(mode == TT_ADAM_SPARSE && grad_arena), "tt_ebc_backward_adam: sparse mode cannot emit a dense grad");
  TT_CHECK_ARG(piece_ws_bytes >= tt_ebc_piece_workspace_bytes(n, E) && (piece_ws || n == 0),
               "tt_ebc_backward_adam: piece workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int G = lanes_per_row(E);
  const int threads = 256, gpb = threads / G;
  GradSrc src{grad0, grad1, ld0, ld1, col0_0, col0_1, n0, K0, K1 > 0 ? K1 : 1, grad_table, n_per_rank > 0 ? n_per_rank : 1, d_colmap0, d_colmap1};
  float* pieces = (float*)piece_ws;
  const int64_t n_chunks = (n + SEG_CHUNK - 1) / SEG_CHUNK;
  const float* gs_buf = (const float*)((const char*)piece_ws + align_up((n_chunks + 1) * 2 * (int64_t)E * 4, 256));
  // Exact-dense mode: the gradients are compacted in sorted order first (independent 16-byte gathers, then streaming sums) — the
  // row kernels walk every arena row and the compacted copy keeps their gradient reads sequential. Sparse mode fetches through
  // the sorted slots directly: the pooled gradients (B*K vectors) sit in the L2 while a compacted copy (B*K*L vectors) would
  // cost an HBM round trip; measured on B200 at the config-5 shard 2.42 ms vs 2.73 ms. That holds for large steps with short
  // segments; a reference-sized step (65 k entries, dozens of entries per row of the smaller tables) is latency-bound and the
  // dependent slot -> gradient chain per entry makes the direct form several times slower there (76 us vs ~10 us), so it is used
  // from 2^19 entries up. Peer (NVLink) sources are always compacted. TT_EBC_DIRECT=0/1 overrides (experiments). Both ways read
  // the same values in the same order: bit-identical.
  static const int direct_env = [] { const char* e = getenv("TT_EBC_DIRECT"); return e ? atoi(e) : -1; }();
  const bool direct = !grad_table && (direct_env >= 0 ? direct_env != 0 : (mode == TT_ADAM_SPARSE && do_adam && n >= ((int64_t)1 << 19)));
  const GradFetch gs{direct ? nullptr : gs_buf, sorted_slots, src, grad_dtype == TT_BF16 ? 1 : 0};
  if (n > 0) {
    if (!direct) {
      int64_t want = ceil_div(n * (E / 4), (int64_t)threads);
      const int64_t cap = (int64_t)sm_count() * 16;
      const int ggrid = (int)(want > cap ? cap : want);
      static const int g4_env = [] { const char* e = getenv("TT_EBC_GATHER4"); return e ? atoi(e) : 1; }();
      const bool local4 = g4_env && !grad_table && (threads % (E / 4)) == 0 && n * (E / 4) < ((int64_t)1 << 30) && col0_0 < ((int64_t)1 << 30) && col0_1 < ((int64_t)1 << 30);
      if (local4) {
        const int g4 = (int)ceil_div(want, 4);
        if (grad_dtype == TT_F32)
          ebc_gather_sorted4_kernel<float><<<g4 < 1 ? 1 : g4, threads, 0, s>>>(sorted_slots, n, d_n, src, E, (float*)gs_buf);
        else
          ebc_gather_sorted4_kernel<__nv_bfloat16><<<g4 < 1 ? 1 : g4, threads, 0, s>>>(sorted_slots, n, d_n, src, E, (float*)gs_buf);
      } else if (grad_dtype == TT_F32)
        ebc_gather_sorted_kernel<float><<<ggrid, threads, 0, s>>>(sorted_slots, n, d_n, src, E, (float*)gs_buf);
      else
        ebc_gather_sorted_kernel<__nv_bfloat16><<<ggrid, threads, 0, s>>>(sorted_slots, n, d_n, src, E, (float*)gs_buf);
    }
    int SUB = 4;
    while (SUB > 1 && (G * SUB > 32 || 32 % (G * SUB) != 0)) SUB >>= 1;
    if (32 % G != 0) SUB = 1;  // lane groups that straddle warps cannot shuffle: one group per chunk
    if (direct)
      ebc_piece_kernel<true><<<ceil_div(n_chunks, threads / (G * SUB)), threads, 0, s>>>(sorted_rows, gs, n, d_n, row_begin, row_end,
                                                                                         E, G, SUB, pieces);
    else
      ebc_piece_kernel<false><<<ceil_div(n_chunks, threads / (G * SUB)), threads, 0, s>>>(sorted_rows, gs, n, d_n, row_begin, row_end, E,
                                                                                          G, SUB, pieces);
  }
  if (mode == TT_ADAM_SPARSE) {
    if (n > 0)
    {
      int64_t want = ceil_div(n, (int64_t)gpb * SP_ROWS);
      const int64_t cap = (int64_t)sm_count() * 12;  // 3 resident CTAs per SM, 4 waves
      auto kern = d_lazy ? (direct ? ebc_rows_sparse_kernel<true, true> : ebc_rows_sparse_kernel<false, true>)
                         : (direct ? ebc_rows_sparse_kernel<true, false> : ebc_rows_sparse_kernel<false, false>);
      kern<<<(int)(want > cap ? cap : want), threads, 0, s>>>(arena_w, arena_m, arena_v, E, G, sorted_rows, gs, n, d_n, row_end, pieces,
                                                              d_hyper, d_lazy);
    }
  } else {
    int64_t want = ceil_div(total_rows, gpb);
    const bool wg = grad_arena != nullptr;
    const bool streaming = do_adam && !wg && total_rows * (int64_t)E * 4 >= ((int64_t)256 << 20);  // arena >> L2
    // L2-resident arena (the reference's sizes): two waves of resident CTAs walking the rows grid-stride instead of one CTA per
    // 32 rows — the kernels launched next to this one (bias folds, split-K reduces) are then placed at once instead of behind
    // its last wave: 275.8 -> 271.2 us per step (three CTAs per SM: 273.3). TT_ROWS_DENSE_CTAS_PER_SM overrides (experiments).
    static const int cap_env = [] { const char* e = getenv("TT_ROWS_DENSE_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    const int64_t cap = (int64_t)sm_count() * (cap_env > 0 ? cap_env : (streaming ? 32 : 6));
    const int grid = (int)(want > cap ? cap : want);
#define TT_ROWS_LAUNCH(A, W)                                                                                         \
  (direct ? ebc_rows_dense_kernel<A, W, true> : ebc_rows_dense_kernel<A, W, false>)<<<grid, threads, 0, s>>>(arena_w, arena_m, arena_v, grad_arena, total_rows, E, G, gs, \
                                                       row_begin, row_end, row_stamp, stamp, d_stamp, pieces, d_hyper, 0)
    if (streaming)
      (direct ? ebc_rows_dense_stream_kernel<true> : ebc_rows_dense_stream_kernel<false>)<<<grid, threads, 0, s>>>(
          arena_w, arena_m, arena_v, total_rows, E, G, gs, row_begin, row_end, row_stamp, stamp, d_stamp, pieces, d_hyper);
    else if (do_adam && wg) TT_ROWS_LAUNCH(true, true);
    else if (do_adam) TT_ROWS_LAUNCH(true, false);
    else TT_ROWS_LAUNCH(false, true);
#undef TT_ROWS_LAUNCH
  }
  TT_LAUNCH_CHECK("ebc_backward_adam");
  return TT_OK;
}

int tt_ebc_backward_adam(float* arena_w, float* arena_m, float* arena_v, float* grad_arena, int64_t total_rows,
                         int32_t E, const int32_t* sorted_rows, const int32_t* sorted_slots, int64_t n,
                         const int32_t* row_begin, const int32_t* row_end, const int32_t* row_stamp, int32_t stamp,
                         const int64_t* d_stamp, const void* grad0, int64_t ld0, int64_t col0_0, int32_t K0, int64_t n0, const void* grad1,
                         int64_t ld1, int64_t col0_1, int32_t K1, int64_t n1, int32_t grad_dtype, void* piece_ws,
                         int64_t piece_ws_bytes, const tt_adam_hyper* d_hyper, int32_t mode, const int32_t* d_n,
                         const void* const* grad_table, int64_t n_per_rank, const int32_t* d_colmap0, const int32_t* d_colmap1,
                         tt_stream_t stream) {
  return backward_adam_impl(arena_w, arena_m, arena_v, grad_arena, total_rows, E, sorted_rows, sorted_slots, n, row_begin, row_end,
                            row_stamp, stamp, d_stamp, grad0, ld0, col0_0, K0, n0, grad1, ld1, col0_1, K1, n1, grad_dtype, piece_ws,
                            piece_ws_bytes, d_hyper, mode, d_n, grad_table, n_per_rank, d_colmap0, d_colmap1, stream, nullptr);
}

/* TT_ADAM_LAZY_EXACT: the touched rows only, each first caught up over the steps it sat out (lazy_replay_vec4) */
int tt_ebc_backward_adam_lazy(float* arena_w, float* arena_m, float* arena_v, int64_t total_rows, int32_t E,
                              const int32_t* sorted_rows, const int32_t* sorted_slots, int64_t n, const int32_t* row_begin,
                              const int32_t* row_end, const int32_t* row_stamp, const int64_t* d_stamp, const void* grad0, int64_t ld0,
                              int64_t col0_0, int32_t K0, int64_t n0, const void* grad1, int64_t ld1, int64_t col0_1, int32_t K1,
                              int64_t n1, int32_t grad_dtype, void* piece_ws, int64_t piece_ws_bytes, const tt_adam_hyper* d_hyper,
                              const int32_t* d_n, const void* const* grad_table, int64_t n_per_rank, const int32_t* d_colmap0,
                              const int32_t* d_colmap1, const tt_lazy_ctx* d_lazy, tt_stream_t stream) {
  TT_CHECK_ARG(d_lazy && arena_w, "tt_ebc_backward_adam_lazy: needs the lazy context and the arena");
  return backward_adam_impl(arena_w, arena_m, arena_v, nullptr, total_rows, E, sorted_rows, sorted_slots, n, row_begin, row_end,
                            row_stamp, 0, d_stamp, grad0, ld0, col0_0, K0, n0, grad1, ld1, col0_1, K1, n1, grad_dtype, piece_ws,
                            piece_ws_bytes, d_hyper, TT_ADAM_SPARSE, d_n, grad_table, n_per_rank, d_colmap0, d_colmap1, stream, d_lazy);
}

int tt_ebc_lazy_flush(const tt_lazy_ctx* d_lazy, int64_t rows, int32_t E, tt_stream_t stream) {
  TT_CHECK_ARG(d_lazy && rows > 0 && E > 0 && (E % 4) == 0, "tt_ebc_lazy_flush: bad arguments");
  const int G = lanes_per_row(E);
  int64_t want = ceil_div(rows, 256 / G);
  const int64_t cap = (int64_t)sm_count() * 32;
  ebc_lazy_flush_kernel<<<(int)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(d_lazy, rows, E, G);
  TT_LAUNCH_CHECK("ebc_lazy_flush_kernel");
  return TT_OK;
}

int64_t tt_ebc_small_workspace_bytes(int32_t B, int64_t small_floats) {
  const int64_t chunks = (B + SM_CHUNK - 1) / SM_CHUNK;
  return chunks * small_floats * 4 + 256;
}

int tt_ebc_small_backward(const int64_t* values, int32_t B, int32_t K, int32_t layout, const void* grad, int64_t ld,
                          int32_t grad_dtype, int32_t E, const void* d_small_tabs, int32_t n_small, int32_t max_vocab,
                          int64_t small_floats, float* gout, void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(values && grad && gout && workspace && B > 0 && K > 0 && E > 0, "tt_ebc_small_backward: bad arguments");
  TT_CHECK_ARG(grad_dtype == TT_F32 || grad_dtype == TT_BF16, "tt_ebc_small_backward: bad dtype");
  if (n_small == 0) return TT_OK;
  TT_CHECK_ARG(d_small_tabs && max_vocab > 0, "tt_ebc_small_backward: missing table descriptors");
  TT_CHECK_ARG(workspace_bytes >= tt_ebc_small_workspace_bytes(B, small_floats), "tt_ebc_small_backward: workspace too small");
  const size_t smem = (size_t)8 * max_vocab * E * sizeof(float);
  TT_CHECK_ARG(smem <= 160 * 1024, "tt_ebc_small_backward: max_vocab * E too large for the shared-memory accumulators");
  cudaStream_t s = (cudaStream_t)stream;
  // ONE wave of CTAs (about four per SM): the fused step launches this kernel next to the large-table update, and the CTAs of a
  // kernel launched behind a multi-wave grid are only placed once that grid's last wave has been — 27 tables x 64 chunks of 128
  // samples (two waves) held the segment-sum kernel back by ~10 us
  const int chunks_max = (B + SM_CHUNK - 1) / SM_CHUNK;
  int chunks = sm_count() * 4 / n_small;
  chunks = chunks < 1 ? 1 : (chunks > chunks_max ? chunks_max : chunks);
  const int chunk = ((B + chunks - 1) / chunks + 63) / 64 * 64;  // whole rounds of the eight warps
  chunks = (B + chunk - 1) / chunk;
  float* part = (float*)workspace;
  dim3 grid(n_small, chunks);
  if (grad_dtype == TT_F32) {
    TT_CUDA(cudaFuncSetAttribute(ebc_small_partial_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ebc_small_partial_kernel<float><<<grid, 256, smem, s>>>(values, B, K, layout, (const float*)grad, ld, E,
                                                            (const SmallTab*)d_small_tabs, part, small_floats, chunk);
  } else {
    TT_CUDA(cudaFuncSetAttribute(ebc_small_partial_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ebc_small_partial_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(values, B, K, layout, (const __nv_bfloat16*)grad, ld, E,
                                                                    (const SmallTab*)d_small_tabs, part, small_floats, chunk);
  }
  ebc_small_reduce_kernel<<<ceil_div(small_floats, 256), 256, 0, s>>>(part, small_floats, chunks, small_floats, gout);
  TT_LAUNCH_CHECK("ebc_small_backward");
  return TT_OK;
}

int tt_dense_adam(float* w, const float* g, float* m, float* v, int64_t n, const tt_adam_hyper* d_hyper,
                  tt_stream_t stream) {
  TT_CHECK_ARG(w && g && m && v && d_hyper && n >= 0, "tt_dense_adam: NULL pointer");
  TT_CHECK_ARG((((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "tt_dense_adam: buffers must be 16-byte aligned");
  if (n == 0) return TT_OK;
  int64_t want = ceil_div(n / 4 + 1, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  dense_adam_kernel<<<(int)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(w, g, m, v, n, d_hyper);
  TT_LAUNCH_CHECK("dense_adam_kernel");
  return TT_OK;
}

}  // extern "C"
