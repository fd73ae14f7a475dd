// On-GPU batch assembly from HBM-resident feature stores.
//
// The reference collates on the CPU: numpy fancy-indexing of dataset.{notice,company}_store['dense_projected'|'categorical']
// by the batch's row indices, a KeyedJaggedTensor built from the flattened [B,K] ids, then an H2D copy of ~1.8 KB per sample
// (src/towers/pairs/unified_bid_data_loader.py:630-684,827-841; scripts/train.py:261-273). With the stores resident in HBM the
// same batch is two row gathers: only the 2 x 8 bytes of indices per sample cross PCIe. Pure byte movement: one warp per
// sample and store, 128-bit loads/stores, output written straight into the train step's static input buffers.
#include "common.cuh"
#include <cuda_fp16.h>

namespace tt {

template <typename T>
__device__ __forceinline__ float4 load4_as_float(const T* p);
template <>
__device__ __forceinline__ float4 load4_as_float<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load4_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(p)));
}
template <>
__device__ __forceinline__ float4 load4_as_float<__half>(const __half* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// dense_out[b, 0:D] = float(dense_store[idx[b], 0:D]);  values_out[slot(b,k)] = int64(cat_store[idx[b], k])
// idx outside [0, n_rows) -> zeros / id 0 and *d_err += 1 (the reference raises IndexError from numpy; a device kernel cannot)
template <typename DT, typename CT>
__global__ void __launch_bounds__(256)
batch_gather_kernel(const DT* __restrict__ dense_store, int64_t dense_ld, int D, const CT* __restrict__ cat_store, int64_t cat_ld,
                    int K, int64_t n_rows, const int64_t* __restrict__ idx, int B, float* __restrict__ dense_out, int64_t out_ld,
                    int64_t* __restrict__ values_out, int layout, int32_t* __restrict__ d_err, int vec) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < B; b += n_warps) {
    int64_t r = __ldg(idx + b);
    const bool ok = r >= 0 && r < n_rows;
    if (!ok) {
      if (lane == 0 && d_err) atomicAdd(d_err, 1);
      r = 0;
    }
    if (dense_out) {
      const DT* src = dense_store + r * dense_ld;
      float* dst = dense_out + (int64_t)b * out_ld;
      const int d4 = vec ? (D & ~3) : 0;  // 128-bit path needs 16-byte aligned output rows
      for (int c = lane * 4; c < d4; c += 128) {
        float4 v = ok ? load4_as_float<DT>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dst + c) = v;
      }
      for (int c = d4 + lane; c < D; c += 32) dst[c] = ok ? (float)src[c] : 0.f;
    }
    if (values_out) {
      const CT* src = cat_store + r * cat_ld;
      for (int k = lane; k < K; k += 32) {
        const int64_t id = ok ? (int64_t)__ldg(src + k) : 0;
        values_out[layout == TT_LAYOUT_SAMPLE_MAJOR ? (int64_t)b * K + k : (int64_t)k * B + b] = id;
      }
    }
  }
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_batch_gather(const void* dense_store, int32_t dense_dtype, int64_t dense_ld, int32_t D, const void* cat_store,
                    int32_t cat_dtype, int64_t cat_ld, int32_t K, int64_t n_rows, const int64_t* idx, int32_t B, float* dense_out,
                    int64_t out_ld, int64_t* values_out, int32_t layout, int32_t* d_err, tt_stream_t stream) {
  TT_CHECK_ARG(idx && B >= 0 && n_rows > 0, "tt_batch_gather: bad idx/B/n_rows");
  TT_CHECK_ARG(!dense_out || (dense_store && D > 0 && dense_ld >= D && out_ld >= D), "tt_batch_gather: bad dense arguments");
  TT_CHECK_ARG(!values_out || (cat_store && K > 0 && cat_ld >= K), "tt_batch_gather: bad categorical arguments");
  TT_CHECK_ARG(dense_out || values_out, "tt_batch_gather: nothing to do");
  TT_CHECK_ARG(dense_dtype == TT_F32 || dense_dtype == TT_BF16 || dense_dtype == TT_F16, "tt_batch_gather: dense dtype must be f32/bf16/f16");
  TT_CHECK_ARG(cat_dtype == TT_I64 || cat_dtype == TT_I32, "tt_batch_gather: categorical dtype must be i64/i32");
  TT_CHECK_ARG(layout == TT_LAYOUT_SAMPLE_MAJOR || layout == TT_LAYOUT_KEY_MAJOR, "tt_batch_gather: bad layout %d", layout);
  int vec = 0;
  if (dense_out) {
    TT_CHECK_ARG(((uintptr_t)dense_store % 16) == 0 && (dense_ld % 4) == 0,
                 "tt_batch_gather: the dense store must be 16-byte aligned with a row pitch that is a multiple of 4 elements");
    vec = ((uintptr_t)dense_out % 16) == 0 && (out_ld % 4) == 0;
  }
  if (B == 0) return TT_OK;
  const int threads = 256;
  int64_t want = ceil_div((int64_t)B * 32, threads);
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want > cap ? cap : want);
  cudaStream_t s = (cudaStream_t)stream;
#define TT_BG(DT, CT)                                                                                                       \
  batch_gather_kernel<DT, CT><<<grid, threads, 0, s>>>((const DT*)dense_store, dense_ld, D, (const CT*)cat_store, cat_ld, K, \
                                                       n_rows, idx, B, dense_out, out_ld, values_out, layout, d_err, vec)
  if (cat_dtype == TT_I64) {
    if (dense_dtype == TT_F32) TT_BG(float, int64_t);
    else if (dense_dtype == TT_BF16) TT_BG(__nv_bfloat16, int64_t);
    else TT_BG(__half, int64_t);
  } else {
    if (dense_dtype == TT_F32) TT_BG(float, int32_t);
    else if (dense_dtype == TT_BF16) TT_BG(__nv_bfloat16, int32_t);
    else TT_BG(__half, int32_t);
  }
#undef TT_BG
  TT_LAUNCH_CHECK("batch_gather_kernel");
  return TT_OK;
}

}  // extern "C"
