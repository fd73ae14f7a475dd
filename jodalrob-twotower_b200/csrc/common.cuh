// Shared device/host helpers for the sm_100a two-tower kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/twotower_b200.h"

namespace tt {

// ---- error plumbing (C ABI: int status + tt_last_error()) -------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TT_CHECK_ARG(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      tt::set_error(__VA_ARGS__);                \
      return TT_ERR_INVALID_ARG;                 \
    }                                            \
  } while (0)

#define TT_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return tt::cuda_fail(e__, #call); \
  } while (0)

#define TT_LAUNCH_CHECK(name)                                      \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return tt::cuda_fail(e__, name);       \
  } while (0)

int sm_count();

// ---- programmatic dependent launch (PDL): kernels of one stream's dependency chain are launched with the
// programmatic-stream-serialization attribute; each of them waits for its predecessor with griddepcontrol.wait AFTER its own set-up
// (barrier init, TMEM allocation) and lets its successor be scheduled early with griddepcontrol.launch_dependents. The predecessor's
// drain + the launch latency (2-3 us per kernel boundary of the latency-bound step) then overlap the successor's prologue.
// Rule: a kernel launched through launch_pdl MUST execute pdl_wait() before it touches global memory (completion is transitive
// along the chain only if every link waits). Measured slower on the multi-stream step (see pdl_enabled): off unless TT_PDL=1 — the
// kernels are then launched as ordinary kernels and the two instructions are no-ops.
bool pdl_enabled();

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void f4_add(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}

__device__ __forceinline__ float4 unpack_bf16x4(const uint2& u) {
  __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

// generic scalar load/store by dtype tag (TT_F32 / TT_BF16)
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// Philox4x32-10 counter RNG: dropout keep-mask is a pure function of (seed, step, stream, element),
// so the backward pass recomputes it instead of storing a mask.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// keep decision for element idx (row-major flat index); 4 elements share one Philox call
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t step, uint32_t stream, uint64_t idx, float p) {
  uint4 c = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), (uint32_t)step, stream ^ (uint32_t)(step >> 32) * 0x9E3779B9u);
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  // uniform in [0,1): keep iff u >= p
  return (float)(w >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace tt
