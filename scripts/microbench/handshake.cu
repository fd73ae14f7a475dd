// Micro-benchmark: cost of the mbarrier / tcgen05.commit handshakes used by the sim-CE pipeline (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../jodalrob-twotower_b200/csrc -o handshake handshake.cu
#include "tc_common.cuh"
#include <cstdio>
using namespace tt;

// mode 0: plain mbarrier ping-pong between two warps (arrive / try_wait)
// mode 1: warp A signals with tcgen05.commit (no MMAs outstanding), warp B answers with mbarrier.arrive
// mode 2: like 1 but 8 "epilogue" warps each arrive once (count 8) before A continues
__global__ void handshake_kernel(int mode, int iters, long long* out) {
  __shared__ uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_b = mode == 2 ? 8 : 1;
  if (threadIdx.x == 0) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_b, n_b);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        if (mode == 0) mbar_arrive(&bar_a); else umma_commit(&bar_a);
        mbar_wait(&bar_b, i & 1);
      }
    }
  } else if (warp <= n_b) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&bar_a, i & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_b);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  for (int mode = 0; mode < 3; ++mode) {
    const int iters = 2000;
    handshake_kernel<<<148, 32 * 9>>>(mode, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mode %d: %s  %.1f cycles per round trip (block 0), %.1f (block 100)\n", mode, cudaGetErrorString(e), (double)h[0] / iters, (double)h[100] / iters);
  }
  return 0;
}
