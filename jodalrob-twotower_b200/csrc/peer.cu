// One-shot all-reduce of a small fp64 vector over NVLink peer memory (no NCCL): the tiny reductions of a sharded train step
// (SyncBN sums, the loss partials, the "all pooled rows have landed" barrier) are latency-bound, and a collective-library call
// costs 15-20 us inside the step's CUDA graph. One single-CTA kernel per rank instead:
//   1. store my vector into slot [parity][my rank] of EVERY rank's mailbox (peer-mapped stores over NVLink),
//   2. __threadfence_system, then release-store the epoch into flag [parity][my rank] of every rank,
//   3. acquire-spin on my own W flags until they show the epoch,
//   4. sum the W slots in rank order (every rank adds the same numbers in the same order: bit-identical results everywhere).
// The epoch is a device-resident counter per call site, so a captured CUDA graph replays correctly; slots alternate by epoch
// parity, and a rank can only be one epoch ahead of the slowest rank at a site (it needs that rank's contribution to finish
// the epoch in between), so a slot is never overwritten before it has been read. The spin is bounded: a missing peer traps
// after ~2 s instead of hanging the GPU.
//
// Kernels of DIFFERENT ranks wait for each other here, so this path is for one-process-per-GPU runs only; the in-process rank
// emulation used by the single-GPU tests keeps the host-mediated collectives (dist.py).
#include "common.cuh"

namespace tt {

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
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 2000000000ull) {
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

}  // namespace tt

using namespace tt;

extern "C" {

int64_t tt_peer_mailbox_bytes(int32_t world, int32_t n_max) {
  if (world <= 0 || n_max <= 0) return 0;
  const int64_t flags = ((int64_t)2 * world + 15) / 16 * 16;
  return (flags + (int64_t)2 * world * n_max) * 8;
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

}  // extern "C"
