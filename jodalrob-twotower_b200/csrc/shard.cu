// Row-wise sharded embedding tables across GPUs (one process per GPU, NVLink peer memory).
//
// Shard map (TorchRec row-wise convention; the reference itself is single-GPU, SURVEY.md §8(c),(e)):
//     block_k = ceil(V_k / world), owner = id / block_k, local row = local_base[k] + id - owner * block_k.
// Per step every rank (1) clamps its ids, (2) all-gathers them (NCCL, 4 B per id), (3) picks out the ids it owns
// with a deterministic stable compaction, (4) PUSHES the looked-up rows straight into the requesters' concat
// buffers through peer-mapped pointers — lookup and the row all-to-all are one kernel — and in the backward
// (5) PULLS the pooled-output gradients of its rows from the requesters' buffers inside the segment-sum kernel
// (ebc.cu, GradSrc::table).  Every buffer has a static size, so the whole step stays CUDA-graph capturable.
#include "common.cuh"
#include "adam.cuh"

namespace tt {

// ids[slot] = clamp(values, 0, V_k-1) as int32, slot = b*K + k (whatever the value layout)
__global__ void __launch_bounds__(256)
clamp_ids_kernel(const int64_t* __restrict__ values, const int32_t* __restrict__ vocab, int B, int K, int layout,
                 int32_t* __restrict__ ids) {
  const int64_t n = (int64_t)B * K;
  for (int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(slot / K), k = (int)(slot - (int64_t)b * K);
    int64_t id = values[layout == TT_LAYOUT_SAMPLE_MAJOR ? slot : (int64_t)k * B + b];
    const int64_t hi = (int64_t)vocab[k] - 1;
    id = id < 0 ? 0 : (id > hi ? hi : id);
    ids[slot] = (int32_t)id;
  }
}

// same for a subset of a tower's keys (the sharded "large" tables): ids[b*KL + kl] = clamp(values(b, keys[kl]))
__global__ void __launch_bounds__(256)
clamp_ids_subset_kernel(const int64_t* __restrict__ values, const int32_t* __restrict__ vocab, const int32_t* __restrict__ keys,
                        int B, int K, int KL, int layout, int32_t* __restrict__ ids) {
  const int64_t n = (int64_t)B * KL;
  for (int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(slot / KL), kl = (int)(slot - (int64_t)b * KL);
    const int k = keys[kl];
    int64_t id = values[layout == TT_LAYOUT_SAMPLE_MAJOR ? (int64_t)b * K + k : (int64_t)k * B + b];
    const int64_t hi = (int64_t)vocab[k] - 1;
    id = id < 0 ? 0 : (id > hi ? hi : id);
    ids[slot] = (int32_t)id;
  }
}

struct RouteMeta {
  const int32_t* ids_all;     // [world * n_per_rank] clamped global ids in slot order, rank-major
  const int32_t* block;       // [K0+K1] rows per rank of table k
  const int32_t* local_base;  // [K0+K1] first local row of table k
  int64_t n_total, n_per_rank, n0;
  int K0, K1, me;
};

__device__ __forceinline__ bool route_one(const RouteMeta& m, int64_t g, int32_t& local_row) {
  const int64_t s = g % m.n_per_rank;
  const int k = s < m.n0 ? (int)(s % m.K0) : m.K0 + (int)((s - m.n0) % m.K1);
  const int32_t id = m.ids_all[g];
  const int32_t blk = m.block[k];
  const int32_t owner = id / blk;
  local_row = m.local_base[k] + id - owner * blk;
  return owner == m.me;
}

constexpr int RT_THREADS = 256, RT_ITEMS = 8, RT_TILE = RT_THREADS * RT_ITEMS;

// pass 1: how many ids of each 2048-entry tile does this rank own
__global__ void __launch_bounds__(RT_THREADS) route_count_kernel(RouteMeta m, int32_t* __restrict__ tile_counts, int n_tiles) {
  __shared__ int32_t s_cnt[RT_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * RT_TILE + (int64_t)warp * (32 * RT_ITEMS);
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < RT_ITEMS; ++i) {
    const int64_t g = base + i * 32 + lane;
    int32_t lr;
    const bool own = g < m.n_total && route_one(m, g, lr);
    cnt += __popc(__ballot_sync(0xffffffffu, own));
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < RT_THREADS / 32; ++w) t += s_cnt[w];
    tile_counts[blockIdx.x] = t;
    if (blockIdx.x == 0) tile_counts[n_tiles] = 0;  // becomes the grand total after the exclusive scan
  }
}

// pass 2 (after an exclusive scan of tile_counts): stable compaction, order == global slot order
__global__ void __launch_bounds__(RT_THREADS)
route_compact_kernel(RouteMeta m, const int32_t* __restrict__ tile_offsets, int n_tiles, int32_t* __restrict__ rows_c,
                     int32_t* __restrict__ slots_c, int32_t* __restrict__ n_owned) {
  __shared__ int32_t s_cnt[RT_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * RT_TILE + (int64_t)warp * (32 * RT_ITEMS);
  bool own[RT_ITEMS];
  int32_t lr[RT_ITEMS];
  int rank_in_warp[RT_ITEMS];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < RT_ITEMS; ++i) {
    const int64_t g = base + i * 32 + lane;
    own[i] = g < m.n_total && route_one(m, g, lr[i]);
    const unsigned b = __ballot_sync(0xffffffffu, own[i]);
    rank_in_warp[i] = cnt + __popc(b & ((1u << lane) - 1u));
    cnt += __popc(b);
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  int warp_off = tile_offsets[blockIdx.x];
  for (int w = 0; w < warp; ++w) warp_off += s_cnt[w];
#pragma unroll
  for (int i = 0; i < RT_ITEMS; ++i) {
    if (own[i]) {
      const int64_t g = base + i * 32 + lane;
      rows_c[warp_off + rank_in_warp[i]] = lr[i];
      slots_c[warp_off + rank_in_warp[i]] = (int32_t)g;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_owned = tile_offsets[n_tiles];
}


// ---------------------------------------------------------------------------------------------------------------------
// All-to-all id routing (the sender buckets, the owner only ever sees its own ids). tt_shard_select_owned above has every rank
// all-gather and scan ALL W * n ids; here each rank
//   1. counts, per 2048-entry tile of its OWN n ids, how many go to each owner            (a2a_count_kernel)
//   2. scans the [owner][tile] counts: a bucket's entries keep their slot order              (scan_exclusive_small_kernel)
//   3. writes (local row, global slot) of every id straight into its owner's receive buffer, segment `sender`, over NVLink
//      peer memory, plus the segment length                                                  (a2a_push_kernel)
//   — barrier —
//   4. the owner packs its W segments (sender-major, slot order inside: exactly the order of the all-gather path, so
//      everything downstream is bit-identical)                                               (a2a_compact_kernel)
// Per rank n * 8 bytes leave over NVLink instead of W * n * 4 arriving, and the per-rank work no longer grows with W.
struct A2aMeta {
  const int32_t* ids;         // [n] this rank's clamped global ids in slot order
  const int32_t* block;       // [K0+K1] rows per rank of table k
  const int32_t* local_base;  // [K0+K1] first local row of table k
  int64_t n, n0;
  int K0, K1, me, world;
};

__device__ __forceinline__ int a2a_route(const A2aMeta& m, int64_t s_, int32_t& local_row) {
  const int k = s_ < m.n0 ? (int)(s_ % m.K0) : m.K0 + (int)((s_ - m.n0) % m.K1);
  const int32_t id = m.ids[s_];
  const int32_t blk = m.block[k];
  const int32_t owner = id / blk;
  local_row = m.local_base[k] + id - owner * blk;
  return owner;
}

constexpr int A2A_MAX_WORLD = 64;

// counts[(owner * (n_tiles + 1)) + tile]; the extra slot per owner stays 0 and becomes the bucket end after the scan
__global__ void __launch_bounds__(RT_THREADS) a2a_count_kernel(A2aMeta m, int32_t* __restrict__ counts, int n_tiles) {
  __shared__ int32_t s_cnt[A2A_MAX_WORLD];
  if (threadIdx.x < m.world) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * RT_TILE + (int64_t)warp * (32 * RT_ITEMS);
#pragma unroll
  for (int i = 0; i < RT_ITEMS; ++i) {
    const int64_t g = base + i * 32 + lane;
    int32_t lr;
    const int owner = g < m.n ? a2a_route(m, g, lr) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, owner);
    if (owner >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&s_cnt[owner], __popc(peers));  // integer adds: order-independent
  }
  __syncthreads();
  if (threadIdx.x < m.world) counts[(int64_t)threadIdx.x * (n_tiles + 1) + blockIdx.x] = s_cnt[threadIdx.x];
  if (blockIdx.x == 0 && threadIdx.x < m.world) counts[(int64_t)threadIdx.x * (n_tiles + 1) + n_tiles] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[(int64_t)m.world * (n_tiles + 1)] = 0;
}

// offsets = exclusive scan of counts (flattened): offsets[o][t] - offsets[o][0] = position of tile t's first id inside bucket o
__global__ void __launch_bounds__(RT_THREADS)
a2a_push_kernel(A2aMeta m, const int32_t* __restrict__ offsets, int n_tiles, int32_t* const* __restrict__ recv_rows,
                int32_t* const* __restrict__ recv_slots, int32_t* const* __restrict__ recv_counts) {
  __shared__ int32_t s_warp[RT_THREADS / 32][A2A_MAX_WORLD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (RT_THREADS / 32) * A2A_MAX_WORLD; i += RT_THREADS) (&s_warp[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RT_TILE + (int64_t)warp * (32 * RT_ITEMS);
  int owner[RT_ITEMS], rank_in_warp[RT_ITEMS];
  int32_t lr[RT_ITEMS];
  // stable rank of every id among the ids of the same owner inside this warp's 256-entry chunk: item i of the warp is entries
  // [32 i, 32 i + 32), so (item, lane) order == slot order
#pragma unroll
  for (int i = 0; i < RT_ITEMS; ++i) {
    const int64_t g = base + i * 32 + lane;
    owner[i] = g < m.n ? a2a_route(m, g, lr[i]) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, owner[i]);
    rank_in_warp[i] = 0;
    if (owner[i] >= 0) {
      rank_in_warp[i] = s_warp[warp][owner[i]] + __popc(peers & ((1u << lane) - 1u));
    }
    __syncwarp();
    if (owner[i] >= 0 && lane == (__ffs(peers) - 1)) s_warp[warp][owner[i]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RT_ITEMS; ++i) {
    if (owner[i] < 0) continue;
    const int o = owner[i];
    int pos = offsets[(int64_t)o * (n_tiles + 1) + blockIdx.x] - offsets[(int64_t)o * (n_tiles + 1)];
    for (int w = 0; w < warp; ++w) pos += s_warp[w][o];
    pos += rank_in_warp[i];
    const int64_t g = base + i * 32 + lane;
    const int64_t dst = (int64_t)m.me * m.n + pos;   // my segment of owner o's receive buffer
    recv_rows[o][dst] = lr[i];
    recv_slots[o][dst] = (int32_t)((int64_t)m.me * m.n + g);  // global slot
  }
  if (blockIdx.x == 0 && threadIdx.x < m.world) {
    const int o = threadIdx.x;
    recv_counts[o][m.me] = offsets[(int64_t)o * (n_tiles + 1) + n_tiles] - offsets[(int64_t)o * (n_tiles + 1)];
  }
}

// owner side: pack the W received segments (counts[sender] entries each) into rows_c / slots_c, sender-major
__global__ void __launch_bounds__(256)
a2a_compact_kernel(const int32_t* __restrict__ recv_rows, const int32_t* __restrict__ recv_slots, const int32_t* __restrict__ counts,
                   int world, int64_t n, int32_t* __restrict__ rows_c, int32_t* __restrict__ slots_c, int32_t* __restrict__ n_owned) {
  __shared__ int32_t s_off[A2A_MAX_WORLD + 1];
  if (threadIdx.x == 0) {
    int run = 0;
    for (int r = 0; r < world; ++r) {
      s_off[r] = run;
      run += counts[r];
    }
    s_off[world] = run;
    if (blockIdx.x == 0) *n_owned = run;
  }
  __syncthreads();
  const int64_t total = (int64_t)world * n;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(j / n);
    const int64_t pos = j - (int64_t)r * n;
    if (pos < counts[r]) {
      rows_c[s_off[r] + pos] = recv_rows[j];
      slots_c[s_off[r] + pos] = recv_slots[j];
    }
  }
}

__global__ void __launch_bounds__(1024) scan_exclusive_small_kernel(int32_t* __restrict__ data, int64_t len) {
  __shared__ int32_t s_part[1024];
  const int t = threadIdx.x;
  const int64_t per = (len + 1023) / 1024;
  const int64_t lo = (int64_t)t * per, hi = (lo + per < len) ? lo + per : len;
  int32_t sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += data[i];
  s_part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    int32_t v = (t >= off) ? s_part[t - off] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int32_t run = s_part[t] - sum;
  for (int64_t i = lo; i < hi; ++i) {
    int32_t c = data[i];
    data[i] = run;
    run += c;
  }
}

// fused lookup + all-to-all: one lane group per owned id gathers the row and stores it into the requester's
// concat buffer (peer-mapped, 128-byte coalesced stores over NVLink)
template <typename OutT>
__global__ void __launch_bounds__(256)
push_rows_kernel(const float* __restrict__ arena, int E, int G, const int32_t* __restrict__ rows_c,
                 const int32_t* __restrict__ slots_c, const int32_t* __restrict__ n_owned, void* const* __restrict__ dst_table,
                 int64_t n_per_rank, int64_t n0, int K0, int K1, int64_t ld0, int64_t ld1, int64_t col0_0, int64_t col0_1,
                 const int32_t* __restrict__ cmap0, const int32_t* __restrict__ cmap1, const tt_lazy_ctx* __restrict__ lazy,
                 int upto_off) {
  const int n = *n_owned;
  tt_lazy_ctx L = {};
  tt_adam_hyper lh = {};
  long long upto = 0;
  if (lazy) {  // lazy exact-replay Adam: the owner catches a row up (read-only) before it sends it
    L = *lazy;
    lh = *L.hyper;
    upto = lh.step + upto_off;
  }
  const int lane = threadIdx.x % G;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  // PR_ROWS owned ids per lane group and iteration. Measured at the config-3 shard (311 k owned ids, 10 M-row tables, N = 2): four
  // rows in flight per group 121 us (89 registers, two CTAs per SM), one row 75 us — occupancy beats per-thread batching here
  constexpr int PR_ROWS = 1;
  for (int64_t e0 = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; e0 < n; e0 += PR_ROWS * groups_total) {
    int32_t row[PR_ROWS];
    OutT* dst[PR_ROWS];
    long long lag_from[PR_ROWS];
#pragma unroll
    for (int u = 0; u < PR_ROWS; ++u) {
      const int64_t e = e0 + (int64_t)u * groups_total;
      row[u] = -1;
      dst[u] = nullptr;
      lag_from[u] = -1;
      if (e < n) {
        row[u] = __ldg(rows_c + e);
        int64_t g = __ldg(slots_c + e);
        const int r = (int)(g / n_per_rank);
        g -= (int64_t)r * n_per_rank;
        if (g < n0) {
          const int b = (int)(g / K0), k = (int)(g - (int64_t)b * K0);
          dst[u] = reinterpret_cast<OutT*>(dst_table[2 * r]) + (int64_t)b * ld0 + (cmap0 ? (int64_t)cmap0[k] : col0_0 + (int64_t)k * E);
        } else {
          g -= n0;
          const int b = (int)(g / K1), k = (int)(g - (int64_t)b * K1);
          dst[u] = reinterpret_cast<OutT*>(dst_table[2 * r + 1]) + (int64_t)b * ld1 + (cmap1 ? (int64_t)cmap1[k] : col0_1 + (int64_t)k * E);
        }
        if (lazy) {
          const long long last = L.row_last[row[u]];
          if (last < upto) lag_from[u] = last + 1;
        }
      }
    }
    for (int c = lane; c < vec_per_row; c += G) {
      float4 v[PR_ROWS];
#pragma unroll
      for (int u = 0; u < PR_ROWS; ++u)
        if (row[u] >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(arena + (int64_t)row[u] * E) + c);
#pragma unroll
      for (int u = 0; u < PR_ROWS; ++u) {
        if (row[u] < 0) continue;
        if (lag_from[u] >= 0) {
          float4 mm = *(reinterpret_cast<const float4*>(L.arena_m + (int64_t)row[u] * E) + c);
          float4 vv = *(reinterpret_cast<const float4*>(L.arena_v + (int64_t)row[u] * E) + c);
          lazy_replay_vec4(v[u], mm, vv, lag_from[u], upto, L, lh);
        }
        if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float4*>(dst[u] + 4 * c) = v[u];
        else *reinterpret_cast<uint2*>(dst[u] + 4 * c) = pack_bf16x4(v[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-hot bags across shards (BASELINE configs[4]: pooling factor L). A bag's L ids live on up to W owners; each owner pre-pools
// ITS ids of the bag and sends one partial sum per (bag, owner) instead of up to L rows ("owners pre-pool their part
// of each bag": the reduce-scatter pattern). The owned entries arrive in global slot order (sender-major, slot order inside), and
// the L id-slots of a bag are consecutive, so a bag's entries at one owner are one run of the list.
//   push_pooled_kernel   one lane group per owned entry; run heads sum their run (ascending slot order) and store the fp32 partial
//                        into the requester's partial buffer [owner][bag][E] over NVLink peer memory
//   sum_partials_kernel  requester: bag = sum over the W owners in rank order (fixed order: deterministic) -> concat buffer
// The backward needs nothing new: every id of a bag pulls the bag's pooled gradient through its slot (GradSrc with one column-map
// entry per (key, position) pseudo-key).
__global__ void __launch_bounds__(256)
push_pooled_kernel(const float* __restrict__ arena, int E, int G, const int32_t* __restrict__ rows_c, const int32_t* __restrict__ slots_c,
                   const int32_t* __restrict__ n_owned, float* const* __restrict__ part_table, int64_t n_per_rank, int L, int64_t n_bags,
                   int me) {
  const int n = *n_owned;
  const int lane = threadIdx.x % G;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; e < n; e += groups_total) {
    const int32_t slot = __ldg(slots_c + e);
    const int32_t bag = slot / L;  // global bag index: n_per_rank is a multiple of L
    if (e > 0 && __ldg(slots_c + e - 1) / L == bag) continue;  // not the head of its run
    int64_t x = e + 1;
    while (x < n && x - e < L && __ldg(slots_c + x) / L == bag) ++x;
    const int r = (int)(slot / n_per_rank);
    const int64_t local_bag = (int64_t)bag - (int64_t)r * (n_per_rank / L);
    float* dst = part_table[r] + ((int64_t)me * n_bags + local_bag) * E;
    for (int c = lane; c < vec_per_row; c += G) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t i = e; i < x; i += 4) {  // up to four rows of the run in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i + u < x) v[u] = __ldg(reinterpret_cast<const float4*>(arena + (int64_t)__ldg(rows_c + i + u) * E) + c);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i + u < x) f4_add(acc, v[u]);
      }
      *reinterpret_cast<float4*>(dst + 4 * c) = acc;
    }
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ part, int world, int64_t n_bags, int E, int G, OutT* __restrict__ c0, int64_t ld0,
                    const int32_t* __restrict__ cmap0, int64_t n0_bags, int K0, OutT* __restrict__ c1, int64_t ld1,
                    const int32_t* __restrict__ cmap1, int K1) {
  const int lane = threadIdx.x % G;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / G);
  const int vec_per_row = E >> 2;
  for (int64_t j = (int64_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G; j < n_bags; j += groups_total) {
    OutT* dst;
    if (j < n0_bags) {
      const int b = (int)(j / K0), k = (int)(j - (int64_t)b * K0);
      dst = c0 + (int64_t)b * ld0 + cmap0[k];
    } else {
      const int64_t jj = j - n0_bags;
      const int b = (int)(jj / K1), k = (int)(jj - (int64_t)b * K1);
      dst = c1 + (int64_t)b * ld1 + cmap1[k];
    }
    for (int c = lane; c < vec_per_row; c += G) {
      float4 acc = *reinterpret_cast<const float4*>(part + j * E + 4 * c);
      for (int w = 1; w < world; ++w) f4_add(acc, *reinterpret_cast<const float4*>(part + ((int64_t)w * n_bags + j) * E + 4 * c));
      if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float4*>(dst + 4 * c) = acc;
      else *reinterpret_cast<uint2*>(dst + 4 * c) = pack_bf16x4(acc);
    }
  }
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_ebc_clamp_ids(const int64_t* values, const int32_t* d_vocab, int32_t B, int32_t K, int32_t layout, int32_t* ids_out,
                     tt_stream_t stream) {
  TT_CHECK_ARG(values && d_vocab && ids_out && B >= 0 && K > 0, "tt_ebc_clamp_ids: bad arguments");
  if (B == 0) return TT_OK;
  int64_t want = ceil_div((int64_t)B * K, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  clamp_ids_kernel<<<(int)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(values, d_vocab, B, K, layout, ids_out);
  TT_LAUNCH_CHECK("clamp_ids_kernel");
  return TT_OK;
}

int tt_ebc_clamp_ids_subset(const int64_t* values, const int32_t* d_vocab, const int32_t* d_keys, int32_t B, int32_t K, int32_t KL,
                            int32_t layout, int32_t* ids_out, tt_stream_t stream) {
  TT_CHECK_ARG(values && d_vocab && d_keys && ids_out && B >= 0 && K > 0 && KL >= 0 && KL <= K, "tt_ebc_clamp_ids_subset: bad arguments");
  if (B == 0 || KL == 0) return TT_OK;
  int64_t want = ceil_div((int64_t)B * KL, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  clamp_ids_subset_kernel<<<(int)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(values, d_vocab, d_keys, B, K, KL, layout,
                                                                                            ids_out);
  TT_LAUNCH_CHECK("clamp_ids_subset_kernel");
  return TT_OK;
}

int64_t tt_shard_workspace_bytes(int64_t n_total) {
  const int64_t tiles = (n_total + RT_TILE - 1) / RT_TILE + 2;
  return tiles * 4 + 256;
}

int tt_shard_select_owned(const int32_t* ids_all, const int32_t* d_block, const int32_t* d_local_base, int64_t n_per_rank,
                          int64_t n0, int32_t K0, int32_t K1, int32_t world, int32_t me, int32_t* rows_out, int32_t* slots_out,
                          int32_t* d_n_owned, void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(ids_all && d_block && d_local_base && rows_out && slots_out && d_n_owned && workspace, "tt_shard_select_owned: NULL pointer");
  TT_CHECK_ARG(world > 0 && me >= 0 && me < world && n_per_rank > 0 && n0 >= 0 && n0 <= n_per_rank && K0 > 0, "tt_shard_select_owned: bad sizes");
  TT_CHECK_ARG(n0 == n_per_rank || K1 > 0, "tt_shard_select_owned: K1 missing");
  const int64_t n_total = n_per_rank * world;
  TT_CHECK_ARG(n_total < (1ll << 31), "tt_shard_select_owned: too many ids for int32 slots");
  TT_CHECK_ARG(workspace_bytes >= tt_shard_workspace_bytes(n_total), "tt_shard_select_owned: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  RouteMeta m{ids_all, d_block, d_local_base, n_total, n_per_rank, n0, K0, K1 > 0 ? K1 : 1, me};
  const int n_tiles = (int)((n_total + RT_TILE - 1) / RT_TILE);
  int32_t* counts = (int32_t*)workspace;
  route_count_kernel<<<n_tiles, RT_THREADS, 0, s>>>(m, counts, n_tiles);
  scan_exclusive_small_kernel<<<1, 1024, 0, s>>>(counts, n_tiles + 1);
  route_compact_kernel<<<n_tiles, RT_THREADS, 0, s>>>(m, counts, n_tiles, rows_out, slots_out, d_n_owned);
  TT_LAUNCH_CHECK("shard_select_owned");
  return TT_OK;
}

int64_t tt_shard_a2a_workspace_bytes(int64_t n_per_rank, int32_t world) {
  const int64_t tiles = (n_per_rank + RT_TILE - 1) / RT_TILE + 1;
  return ((int64_t)world * tiles + 2) * 4 + 256;
}

int tt_shard_a2a_send(const int32_t* ids_local, const int32_t* d_block, const int32_t* d_local_base, int64_t n_per_rank, int64_t n0,
                      int32_t K0, int32_t K1, int32_t world, int32_t me, void* const* d_recv_rows, void* const* d_recv_slots,
                      void* const* d_recv_counts, void* workspace, int64_t workspace_bytes, tt_stream_t stream) {
  TT_CHECK_ARG(ids_local && d_block && d_local_base && d_recv_rows && d_recv_slots && d_recv_counts && workspace, "tt_shard_a2a_send: NULL pointer");
  TT_CHECK_ARG(world > 0 && world <= A2A_MAX_WORLD && me >= 0 && me < world && n_per_rank > 0 && n0 >= 0 && n0 <= n_per_rank && K0 > 0,
               "tt_shard_a2a_send: bad sizes");
  TT_CHECK_ARG(n0 == n_per_rank || K1 > 0, "tt_shard_a2a_send: K1 missing");
  TT_CHECK_ARG(n_per_rank * world < (1ll << 31), "tt_shard_a2a_send: too many ids for int32 slots");
  TT_CHECK_ARG(workspace_bytes >= tt_shard_a2a_workspace_bytes(n_per_rank, world), "tt_shard_a2a_send: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  A2aMeta m{ids_local, d_block, d_local_base, n_per_rank, n0, K0, K1 > 0 ? K1 : 1, me, world};
  const int n_tiles = (int)((n_per_rank + RT_TILE - 1) / RT_TILE);
  int32_t* counts = (int32_t*)workspace;
  a2a_count_kernel<<<n_tiles, RT_THREADS, 0, s>>>(m, counts, n_tiles);
  scan_exclusive_small_kernel<<<1, 1024, 0, s>>>(counts, (int64_t)world * (n_tiles + 1) + 1);
  a2a_push_kernel<<<n_tiles, RT_THREADS, 0, s>>>(m, counts, n_tiles, (int32_t* const*)d_recv_rows, (int32_t* const*)d_recv_slots,
                                                 (int32_t* const*)d_recv_counts);
  TT_LAUNCH_CHECK("shard_a2a_send");
  return TT_OK;
}

int tt_shard_a2a_compact(const int32_t* recv_rows, const int32_t* recv_slots, const int32_t* recv_counts, int32_t world,
                         int64_t n_per_rank, int32_t* rows_out, int32_t* slots_out, int32_t* d_n_owned, tt_stream_t stream) {
  TT_CHECK_ARG(recv_rows && recv_slots && recv_counts && rows_out && slots_out && d_n_owned && world > 0 && world <= A2A_MAX_WORLD &&
               n_per_rank > 0, "tt_shard_a2a_compact: bad arguments");
  int64_t want = ceil_div((int64_t)world * n_per_rank, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  a2a_compact_kernel<<<(int)(want > cap ? cap : want), 256, 0, (cudaStream_t)stream>>>(recv_rows, recv_slots, recv_counts, world, n_per_rank,
                                                                                       rows_out, slots_out, d_n_owned);
  TT_LAUNCH_CHECK("a2a_compact_kernel");
  return TT_OK;
}

int tt_ebc_push_pooled(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                       int64_t capacity, void* const* part_table, int64_t n_per_rank, int32_t L, int64_t n_bags, int32_t me,
                       tt_stream_t stream) {
  TT_CHECK_ARG(arena && rows && slots && d_n_owned && part_table && E > 0 && (E % 4) == 0 && L > 0 && n_bags > 0 && n_per_rank == n_bags * L,
               "tt_ebc_push_pooled: bad arguments");
  int g = 1;
  while (g < (E >> 2) && g < 32) g <<= 1;
  int64_t want = ceil_div(capacity, 256 / g);
  const int64_t cap = (int64_t)sm_count() * 16;
  push_pooled_kernel<<<(int)(want < 1 ? 1 : (want > cap ? cap : want)), 256, 0, (cudaStream_t)stream>>>(
      arena, E, g, rows, slots, d_n_owned, (float* const*)part_table, n_per_rank, L, n_bags, me);
  TT_LAUNCH_CHECK("push_pooled_kernel");
  return TT_OK;
}

int tt_ebc_sum_partials(const float* part, int32_t world, int64_t n_bags, int32_t E, void* c0, int64_t ld0, const int32_t* d_colmap0,
                        int64_t n0_bags, int32_t K0, void* c1, int64_t ld1, const int32_t* d_colmap1, int32_t K1, int32_t out_dtype,
                        tt_stream_t stream) {
  TT_CHECK_ARG(part && world > 0 && n_bags > 0 && E > 0 && (E % 4) == 0 && c0 && d_colmap0 && K0 > 0 && n0_bags >= 0 && n0_bags <= n_bags,
               "tt_ebc_sum_partials: bad arguments");
  TT_CHECK_ARG(n0_bags == n_bags || (c1 && d_colmap1 && K1 > 0), "tt_ebc_sum_partials: second tower missing");
  TT_CHECK_ARG(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_ebc_sum_partials: bad dtype");
  int g = 1;
  while (g < (E >> 2) && g < 32) g <<= 1;
  int64_t want = ceil_div(n_bags, 256 / g);
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(want > cap ? cap : want);
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == TT_F32)
    sum_partials_kernel<float><<<grid, 256, 0, s>>>(part, world, n_bags, E, g, (float*)c0, ld0, d_colmap0, n0_bags, K0, (float*)c1, ld1,
                                                    d_colmap1, K1 > 0 ? K1 : 1);
  else
    sum_partials_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(part, world, n_bags, E, g, (__nv_bfloat16*)c0, ld0, d_colmap0, n0_bags, K0,
                                                            (__nv_bfloat16*)c1, ld1, d_colmap1, K1 > 0 ? K1 : 1);
  TT_LAUNCH_CHECK("sum_partials_kernel");
  return TT_OK;
}

int tt_zero_bytes(void* p, int64_t bytes, tt_stream_t stream) {
  TT_CHECK_ARG(p && bytes >= 0, "tt_zero_bytes: bad arguments");
  TT_CUDA(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
  return TT_OK;
}

static int push_rows_impl(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                          int64_t capacity, void* const* dst_table, int64_t n_per_rank, int64_t n0, int32_t K0, int32_t K1, int64_t ld0,
                          int64_t ld1, int64_t col0_0, int64_t col0_1, const int32_t* d_colmap0, const int32_t* d_colmap1,
                          int32_t out_dtype, tt_stream_t stream, const tt_lazy_ctx* d_lazy, int upto_off) {
  TT_CHECK_ARG(arena && rows && slots && d_n_owned && dst_table && E > 0 && (E % 4) == 0, "tt_ebc_push_rows: bad arguments");
  TT_CHECK_ARG(out_dtype == TT_F32 || out_dtype == TT_BF16, "tt_ebc_push_rows: bad dtype");
  int g = 1;
  while (g < (E >> 2) && g < 32) g <<= 1;
  int64_t want = ceil_div(capacity, 256 / g);
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == TT_F32)
    push_rows_kernel<float><<<grid, 256, 0, s>>>(arena, E, g, rows, slots, d_n_owned, dst_table, n_per_rank, n0, K0, K1 > 0 ? K1 : 1,
                                                 ld0, ld1, col0_0, col0_1, d_colmap0, d_colmap1, d_lazy, upto_off);
  else
    push_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(arena, E, g, rows, slots, d_n_owned, dst_table, n_per_rank, n0, K0,
                                                         K1 > 0 ? K1 : 1, ld0, ld1, col0_0, col0_1, d_colmap0, d_colmap1, d_lazy,
                                                         upto_off);
  TT_LAUNCH_CHECK("push_rows_kernel");
  return TT_OK;
}

int tt_ebc_push_rows(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                     int64_t capacity, void* const* dst_table, int64_t n_per_rank, int64_t n0, int32_t K0, int32_t K1, int64_t ld0,
                     int64_t ld1, int64_t col0_0, int64_t col0_1, const int32_t* d_colmap0, const int32_t* d_colmap1,
                     int32_t out_dtype, tt_stream_t stream) {
  return push_rows_impl(arena, E, rows, slots, d_n_owned, capacity, dst_table, n_per_rank, n0, K0, K1, ld0, ld1, col0_0, col0_1,
                        d_colmap0, d_colmap1, out_dtype, stream, nullptr, 0);
}

int tt_ebc_push_rows_lazy(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                          int64_t capacity, void* const* dst_table, int64_t n_per_rank, int64_t n0, int32_t K0, int32_t K1, int64_t ld0,
                          int64_t ld1, int64_t col0_0, int64_t col0_1, const int32_t* d_colmap0, const int32_t* d_colmap1,
                          int32_t out_dtype, const tt_lazy_ctx* d_lazy, int32_t upto_offset, tt_stream_t stream) {
  TT_CHECK_ARG(d_lazy && (upto_offset == 0 || upto_offset == -1), "tt_ebc_push_rows_lazy: bad arguments");
  return push_rows_impl(arena, E, rows, slots, d_n_owned, capacity, dst_table, n_per_rank, n0, K0, K1, ld0, ld1, col0_0, col0_1,
                        d_colmap0, d_colmap1, out_dtype, stream, d_lazy, upto_offset);
}

}  // extern "C"
