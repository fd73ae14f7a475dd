// One-shot all-reduce of a small fp64 vector over NVLink peer memory (no NCCL): the tiny reductions of a sharded train step
// (SyncBN sums, the loss partials, the "all pooled rows have landed" barrier) are latency-bound, and a collective-library call
// costs 15-20 us inside the step's CUDA graph. One single-CTA kernel per rank instead:
//   1. store my vector into slot [parity][my rank] of EVERY rank's mailbox (peer-mapped stores over NVLink),
//   2. __threadfence_system, then release-store the epoch into flag [parity][my rank] of every rank,
//   3. acquire-spin on my own W flags until they show the epoch,
//   4. sum the W slots in rank order (every rank adds the same numbers in the same order: bit-identical results everywhere).
// The epoch is a device-resident counter per call site, so a captured CUDA graph replays correctly; slots alternate by epoch
// parity, and a rank can only be one epoch ahead of the slowest rank at a site (it needs that rank's contribution to finish
// the epoch in between), so a slot is never overwritten before it has been read. The spin is bounded by a configurable
// watchdog (tt_peer_set_timeout_ms; default 10 minutes, 0 = wait for ever like NCCL without a watchdog): legitimate rank skew
// — a checkpoint written by rank 0, a data-loader stall — is waited out, a dead peer traps instead of hanging the GPU for ever.
//
// Kernels of DIFFERENT ranks wait for each other here, so this path is for one-process-per-GPU runs only; the in-process rank
// emulation used by the single-GPU tests keeps the host-mediated collectives (dist.py).
#include "common.cuh"

namespace tt {

__device__ unsigned long long g_peer_timeout_ns = 600ull * 1000000000ull;  // 0 = no watchdog

// mailbox of one site: flags u64 [2][world] | pad to 128 B | slots f64 [2][world][n_max]
__device__ __forceinline__ size_t site_flag_off(int par, int r, int world) { return (size_t)par * world + r; }
__device__ __forceinline__ size_t site_slot_off(int par, int r, int world, int n_max) {
  const size_t flags = ((size_t)2 * world + 15) / 16 * 16;  // u64 units, 128-byte aligned
  return flags + ((size_t)par * world + r) * n_max;
}

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* __restrict__ data, int n, int rank, int world, unsigned long long* const* __restrict__ mailboxes,
                          int n_max, unsigned long long* __restrict__ d_epoch) {
  const unsigned long long epoch = *d_epoch + 1;
  const int par = (int)(epoch & 1);
  const int t = threadIdx.x;
  // 1. my contribution into everybody's mailbox
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(mailboxes[p] + site_slot_off(par, rank, world, n_max));
    for (int i = t; i < n; i += blockDim.x) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait for everybody
  if (t < world) {
    unsigned long long* flag = mailboxes[t] + site_flag_off(par, rank, world);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
    const unsigned long long* mine = mailboxes[rank] + site_flag_off(par, t, world);
    unsigned long long seen = 0, t0 = 0, now = 0;
    const unsigned long long timeout_ns = g_peer_timeout_ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (timeout_ns && now - t0 > timeout_ns) {
        printf("tt_peer_allreduce: rank %d timed out waiting for rank %d (epoch %llu, saw %llu)\n", rank, t, epoch, seen);
        __trap();
      }
    }
  }
  __syncthreads();
  // 4. fixed-order sum
  const double* slots = reinterpret_cast<const double*>(mailboxes[rank] + site_slot_off(par, 0, world, n_max));
  for (int i = t; i < n; i += blockDim.x) {
    double a = slots[i];
    for (int r = 1; r < world; ++r) a += slots[(size_t)r * n_max + i];
    data[i] = a;
  }
  if (t == 0) *d_epoch = epoch;
}

// All-gather over NVLink peer memory (no NCCL): every rank stores its block into EVERY rank's destination buffer at its own offset
// (128-bit peer stores, all CTAs), and the kernel returns only when every rank's block has landed in MY buffer:
//   each CTA: copy its share to all W destinations, __threadfence_system, count itself done;
//   the last CTA to finish: release-store the epoch into flag [parity][my rank] of every rank's mailbox, acquire-spin on my own
//   W flags, reset the counter, bump the epoch.
// Used for the id exchange, the embedding all-gathers for global negatives and the row scales of the sharded step, where an
// NCCL all-gather costs 11-16 us of kernel plus 3-13 us of launch gap for 0.06-1 MB. Reuse of a destination buffer across steps
// is safe for the same reason as the mailboxes: a rank can only be one exchange ahead of the slowest rank at a site.
__global__ void __launch_bounds__(256)
peer_allgather_kernel(const uint4* __restrict__ src, int64_t n16, int rank, int world, uint4* const* __restrict__ dst_bases,
                      int64_t dst_off16, unsigned long long* const* __restrict__ mailboxes, unsigned long long* __restrict__ d_epoch,
                      unsigned int* __restrict__ d_count) {
  __shared__ int s_last;
  const unsigned long long epoch = *d_epoch + 1;  // every CTA reads it before the last one can have bumped it
  const int par = (int)(epoch & 1);
  const int t = threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + t; i < n16; i += stride) {
    const uint4 v = __ldg(src + i);
    for (int p = 0; p < world; ++p) dst_bases[p][dst_off16 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(d_count, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();  // the other CTAs' stores (fenced before their count) are ordered before the release below
  if (t < world) {
    unsigned long long* flag = mailboxes[t] + site_flag_off(par, rank, world);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
    const unsigned long long* mine = mailboxes[rank] + site_flag_off(par, t, world);
    unsigned long long seen = 0, t0 = 0, now = 0;
    const unsigned long long timeout_ns = g_peer_timeout_ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (timeout_ns && now - t0 > timeout_ns) {
        printf("tt_peer_allgather: rank %d timed out waiting for rank %d (epoch %llu, saw %llu)\n", rank, t, epoch, seen);
        __trap();
      }
    }
  }
  __syncthreads();
  if (t == 0) {
    *d_count = 0;
    *d_epoch = epoch;
  }
}

// out[i] = sum_r slots[r][i] in rank order (bit-identical on every rank): the reduce half of an all-reduce whose exchange half was
// a peer all-gather
// Every slot may carry a tail of n_tail values behind the n summed ones: recip_out[r * n_tail + i] = 1 / slots[r][n + i] (the
// row scales of the logits travel with the column-sum partials in one exchange).
__global__ void __launch_bounds__(256)
sum_slots_f32_kernel(const float* __restrict__ slots, int world, int64_t n, int64_t n_tail, float* __restrict__ out,
                     float* __restrict__ recip_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pitch = n + n_tail;
  if (i < n) {
    float a = slots[i];
    for (int r = 1; r < world; ++r) a += slots[(int64_t)r * pitch + i];
    out[i] = a;
  } else if (i < n + (int64_t)world * n_tail) {
    const int64_t k = i - n, r = k / n_tail, j = k - r * n_tail;
    recip_out[k] = 1.f / slots[r * pitch + n + j];
  }
}

}  // namespace tt

using namespace tt;

extern "C" {

int64_t tt_peer_mailbox_bytes(int32_t world, int32_t n_max) {
  if (world <= 0 || n_max <= 0) return 0;
  const int64_t flags = ((int64_t)2 * world + 15) / 16 * 16;
  return (flags + (int64_t)2 * world * n_max) * 8;
}

int tt_peer_set_timeout_ms(int64_t ms) {
  TT_CHECK_ARG(ms >= 0, "tt_peer_set_timeout_ms: negative timeout");
  const unsigned long long ns = (unsigned long long)ms * 1000000ull;
  TT_CUDA(cudaMemcpyToSymbol(g_peer_timeout_ns, &ns, sizeof(ns)));
  return TT_OK;
}

int tt_peer_allreduce_f64(double* data, int32_t n, int32_t rank, int32_t world, const void* const* d_mailboxes, int32_t n_max,
                          int64_t* d_epoch, tt_stream_t stream) {
  TT_CHECK_ARG(data && d_mailboxes && d_epoch && world > 0 && world <= 64 && rank >= 0 && rank < world,
               "tt_peer_allreduce_f64: bad arguments");
  TT_CHECK_ARG(n >= 0 && n <= n_max, "tt_peer_allreduce_f64: n (%d) exceeds the mailbox slot size (%d)", n, n_max);
  peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(data, n, rank, world,
                                                                 (unsigned long long* const*)d_mailboxes, n_max,
                                                                 reinterpret_cast<unsigned long long*>(d_epoch));
  TT_LAUNCH_CHECK("peer_allreduce_f64_kernel");
  return TT_OK;
}

int tt_peer_allgather(const void* src, int64_t bytes, int32_t rank, int32_t world, void* const* d_dst_bases, int64_t dst_offset_bytes,
                      const void* const* d_mailboxes, int64_t* d_epoch, int32_t* d_count, tt_stream_t stream) {
  TT_CHECK_ARG(src && d_dst_bases && d_mailboxes && d_epoch && d_count && world > 0 && world <= 64 && rank >= 0 && rank < world,
               "tt_peer_allgather: bad arguments");
  TT_CHECK_ARG(bytes > 0 && (bytes % 16) == 0 && (dst_offset_bytes % 16) == 0 && ((uintptr_t)src % 16) == 0,
               "tt_peer_allgather: block size, offset and source must be multiples of 16 bytes");
  const int64_t n16 = bytes / 16;
  int grid = (int)((n16 + 255) / 256);
  if (grid > 64) grid = 64;
  peer_allgather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, n16, rank, world, (uint4* const*)d_dst_bases,
                                                                dst_offset_bytes / 16, (unsigned long long* const*)d_mailboxes,
                                                                reinterpret_cast<unsigned long long*>(d_epoch),
                                                                reinterpret_cast<unsigned int*>(d_count));
  TT_LAUNCH_CHECK("peer_allgather_kernel");
  return TT_OK;
}

int tt_sum_slots_f32(const float* slots, int32_t world, int64_t n, int64_t n_tail, float* out, float* recip_out, tt_stream_t stream) {
  TT_CHECK_ARG(slots && out && world > 0 && n > 0 && n_tail >= 0 && (n_tail == 0 || recip_out), "tt_sum_slots_f32: bad arguments");
  const int64_t total = n + (int64_t)world * n_tail;
  sum_slots_f32_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(slots, world, n, n_tail, out, recip_out);
  TT_LAUNCH_CHECK("sum_slots_f32_kernel");
  return TT_OK;
}

}  // extern "C"
