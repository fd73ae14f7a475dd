// Error plumbing, device queries and the device-resident optimizer bookkeeping.
#include "common.cuh"
#include <stdlib.h>
#include <math.h>
#include <string.h>

namespace tt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return TT_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// t += 1; lr_t from the LambdaLR warm-up; bias corrections in double like torch's python scalars
// (torch/optim/adam.py _single_tensor_adam: bias_correction1 = 1 - beta1 ** step, step_size = lr / bc1,
//  bias_correction2_sqrt = sqrt(1 - beta2 ** step)).
__global__ void hyper_advance_kernel(tt_adam_hyper* h, float base_lr, long long warmup, float b1, float b2, float eps,
                                     float wd, unsigned long long seed, float* hist, long long hist_len) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long t = h->step + 1;
  long long s = h->sched_step;
  double lr = (double)base_lr;
  if (warmup > 0 && s < warmup) lr = (double)base_lr * ((double)s / (double)warmup);
  double bc1 = 1.0 - pow((double)b1, (double)t);
  double bc2 = 1.0 - pow((double)b2, (double)t);
  h->step = t;
  h->sched_step = s + 1;
  h->lr = (float)lr;
  h->step_size = (float)(lr / bc1);
  h->bc2_sqrt = (float)sqrt(bc2);
  h->beta1 = b1;
  h->beta2 = b2;
  h->eps = eps;
  h->weight_decay = wd;
  h->rng_seed = seed;
  h->rng_step = (unsigned long long)t;
  if (hist) {  // lazy exact-replay Adam: the per-step scalars a skipped step is replayed with
    hist[2 * (t & (hist_len - 1))] = h->step_size;
    hist[2 * (t & (hist_len - 1)) + 1] = h->bc2_sqrt;
  }
}

bool pdl_enabled() {
  // Measured at the README-default step (B = 8192, bf16, one B200): 331.5 us per step with the attribute, 310.7 us without — the
  // early-scheduled successor CTAs sit on SM slots that the other streams' kernels (second tower, sort, weight gradients) would
  // have used. Off unless TT_PDL=1.
  static const bool on = [] {
    const char* e = getenv("TT_PDL");
    return e && e[0] == '1';
  }();
  return on;
}

}  // namespace tt

extern "C" {

const char* tt_last_error(void) { return tt::g_err; }
int tt_version(void) { return TT_ABI_VERSION; }
int tt_device_sm_count(void) { return tt::sm_count(); }

int tt_hyper_advance(tt_adam_hyper* d_hyper, float base_lr, int64_t warmup_steps, float beta1, float beta2, float eps,
                     float weight_decay, uint64_t rng_seed, tt_stream_t stream) {
  TT_CHECK_ARG(d_hyper != nullptr, "tt_hyper_advance: d_hyper is NULL");
  tt::hyper_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_hyper, base_lr, (long long)warmup_steps, beta1, beta2,
                                                              eps, weight_decay, (unsigned long long)rng_seed, nullptr, 0);
  TT_LAUNCH_CHECK("hyper_advance_kernel");
  return TT_OK;
}

int tt_hyper_advance_hist(tt_adam_hyper* d_hyper, float base_lr, int64_t warmup_steps, float beta1, float beta2, float eps,
                          float weight_decay, uint64_t rng_seed, float* d_hist, int64_t hist_len, tt_stream_t stream) {
  TT_CHECK_ARG(d_hyper != nullptr && d_hist != nullptr && hist_len > 0 && (hist_len & (hist_len - 1)) == 0,
               "tt_hyper_advance_hist: needs a history ring with a power-of-two length");
  tt::hyper_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_hyper, base_lr, (long long)warmup_steps, beta1, beta2,
                                                              eps, weight_decay, (unsigned long long)rng_seed, d_hist,
                                                              (long long)hist_len);
  TT_LAUNCH_CHECK("hyper_advance_kernel");
  return TT_OK;
}

}  // extern "C"
