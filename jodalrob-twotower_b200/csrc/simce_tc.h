// tcgen05 sim-CE entry points (simce_tc.cu), used by the C-ABI wrappers in simce.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace tt {
int64_t simce_tc_workspace_bytes(int64_t Bx, int64_t By, int F);
int simce_tc_rows(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                  float* row_sumexp, float* row_sum_s, float* diag, int32_t* row_argmax, int want_metrics, void* workspace,
                  cudaStream_t s, int sums_only = 0, float* out_inv = nullptr, int pre = 0);
int simce_tc_rows_metrics(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, float* row_sum_s,
                          int32_t* row_argmax, void* workspace, cudaStream_t s);
int simce_tc_bwd(const void* X, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                 const float* inv_x, const float* inv_y, float gscale, float tgt_diag, float tgt_off, float* dX, void* workspace,
                 cudaStream_t s);
// three-pass train step on pre-scaled operands (see simce_tc.cu)
int64_t simce_tc_grad_workspace_bytes(int64_t Bx, int64_t By, int F);
int simce_tc_rows_pre(const void* Xs, const void* Y, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset, float* diag,
                      void* workspace, cudaStream_t s);
int simce_tc_rows_finish(const void* workspace, int64_t Bx, int64_t By, int F, float inv_temp, float* row_sumexp, float* inv_row,
                         const void* X_unit, float gscale, void* Xs_out, cudaStream_t s);
int simce_tc_grad_pre(const void* Xs, const void* Y, const void* Ys, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset,
                      const float* inv_x, float gscale, int smooth, void* workspace, cudaStream_t s);
int simce_tc_grad_finish(const void* workspace, int64_t Bx, int64_t By, int F, float inv_temp, int64_t diag_offset, const void* Y,
                         int deferred, float gscale, float tgt_diag, float tgt_off, float* dX, float* sumexp_out, float* inv_out,
                         const void* X_unit, void* Xs_out, cudaStream_t s);
int simce_tc_scale_rows(const void* Y_unit, const float* sumexp, float gscale, float inv_temp, int64_t By, int F, void* Ys_out,
                        cudaStream_t s);
void simce_tc_set_force_manual(int v);
void simce_tc_set_debug(int v);
void simce_tc_set_timeline(long long* buf);
}  // namespace tt
