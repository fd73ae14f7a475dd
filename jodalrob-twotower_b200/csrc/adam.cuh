// Adam element update with a pinned instruction sequence, and the lazy exact-replay of skipped steps (shared by the table kernels).
#pragma once
#include "common.cuh"

namespace tt {

// torch.optim.Adam single-tensor update, element-wise (op order of torch/optim/adam.py). Every operation is pinned with an
// intrinsic, so the instruction sequence — and with it every bit of the result — is the same in whichever kernel this is
// inlined (no context-dependent FMA contraction): the lazy replay below relies on that to be bit-identical to the dense update.
// The square root and the two divisions use the hardware approximations (sqrt.approx / div.approx: <= 2 ulp, one MUFU operation
// each instead of a ~20-instruction IEEE sequence): the update term they feed is ~lr relative to the weight, so the deviation from
// the IEEE result is ~1e-7 * lr — five orders of magnitude inside the 1e-5 parity bound — and the lazy replay, which is pure
// arithmetic (rows x E element-steps per step in steady state), runs ~3x faster. TT_ADAM_IEEE=1 at compile time restores IEEE ops.
__device__ __forceinline__ float adam_sqrt(float x) {
#ifdef TT_ADAM_IEEE
  return __fsqrt_rn(x);
#else
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ float adam_div(float a, float b) {
#ifdef TT_ADAM_IEEE
  return __fdiv_rn(a, b);
#else
  float y;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b));
  return y;
#endif
}
__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, const tt_adam_hyper& h) {
  g = __fmaf_rn(h.weight_decay, w, g);                                     // grad.add(param, alpha=weight_decay)
  m = __fmaf_rn(__fsub_rn(1.0f, h.beta1), __fsub_rn(g, m), m);             // exp_avg.lerp_(grad, 1 - beta1)
  v = __fmaf_rn(__fmul_rn(__fsub_rn(1.0f, h.beta2), g), g, __fmul_rn(v, h.beta2));  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  const float denom = __fadd_rn(adam_div(adam_sqrt(v), h.bc2_sqrt), h.eps);
  w = __fmaf_rn(-h.step_size, adam_div(m, denom), w);                      // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__device__ __forceinline__ void adam_vec4(float4& w, float4& m, float4& v, const float4& g, const tt_adam_hyper& h) {
  adam_elem(w.x, m.x, v.x, g.x, h);
  adam_elem(w.y, m.y, v.y, g.y, h);
  adam_elem(w.z, m.z, v.z, g.z, h);
  adam_elem(w.w, m.w, v.w, g.w, h);
}

// Lazy exact-replay Adam (TT_ADAM_LAZY_EXACT). The reference's optimizer is DENSE: every embedding row moves every step, touched or
// not (g = wd * w, scripts/train.py:231). Streaming every row every step is what that costs at 10 M rows per table; instead a row's
// (w, m, v) stay untouched in HBM until the row is read or updated again, together with the step `row_last` they are current to, and
// the skipped steps last+1 .. upto — zero batch gradient, each with ITS OWN step size / bias correction from the history ring —
// are replayed in registers then. Same operations in the same order as the dense update: bit-identical results, no HBM traffic
// for rows a batch does not touch.
__device__ __forceinline__ void lazy_replay_vec4(float4& w, float4& m, float4& v, long long from, long long upto, const tt_lazy_ctx& L,
                                                 tt_adam_hyper h) {
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float2* hist = reinterpret_cast<const float2*>(L.hist);
  const long long mask = L.hist_len - 1;
  for (long long s_ = from; s_ <= upto; ++s_) {
    const float2 hs = __ldg(hist + (s_ & mask));
    h.step_size = hs.x;
    h.bc2_sqrt = hs.y;
    adam_vec4(w, m, v, zero, h);
  }
}

}  // namespace tt
