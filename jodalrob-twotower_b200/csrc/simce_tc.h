// tcgen05 sim-CE entry points (simce_tc.cu), used by the C-ABI wrappers in simce.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace tt {
int64_t simce_tc_workspace_bytes(int64_t Bx, int64_t By, int F);
int simce_tc_rows(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                  float* row_sumexp, float* row_sum_s, float* diag, int32_t* row_argmax, int want_metrics, void* workspace,
                  cudaStream_t s, int sums_only = 0, float* out_inv = nullptr);
int simce_tc_rows_metrics(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, float* row_sum_s,
                          int32_t* row_argmax, void* workspace, cudaStream_t s);
int simce_tc_bwd(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                 const float* inv_x, const float* inv_y, float gscale, float tgt_diag, float tgt_off, float* dX, void* workspace,
                 cudaStream_t s);
void simce_tc_set_force_manual(int v);
void simce_tc_set_debug(int v);
void simce_tc_set_timeline(long long* buf);
}  // namespace tt
