"""TEST INFRASTRUCTURE ONLY — comparison helpers for updated weights.

Adam's update lr * m_hat / (sqrt(v_hat) + eps) is ill-conditioned where an element's gradient is ~eps (1e-8):
an absolute gradient difference of 1e-9 (fp32 summation order) then moves the update by a sizeable fraction of lr.
Any two fp32 implementations — the reference on CPU vs the reference on GPU included — disagree on a handful of
such elements by up to ~lr while agreeing to 1e-6 everywhere else.  So "updated weights within 1e-5" is checked as
  (1) without an exact answer at hand: at most max(2, max_bad_frac * numel) of a tensor's elements differ from the fp32
      reference by more than rtol * max|ref|;
  (2) against the fp64 explicit-math oracle when available (the stronger statement, used instead of (1)): the number of our elements further than rtol from the
      exact trajectory is at most 2 + 3x the number of such elements of the fp32 reference itself, and our worst
      element is within an order of magnitude of the reference's worst.
"""
from __future__ import annotations

import torch


def rel_to_max(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def bad_fraction(a: torch.Tensor, b: torch.Tensor, rtol: float) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).abs() > rtol * (b.abs().max() + 1e-30)).double().mean())


def assert_updated_close(mine, ref32, truth64=None, rtol=1e-5, max_bad_frac=1e-3, name=""):
    n = mine.numel()
    if truth64 is None:
        bf = bad_fraction(mine, ref32, rtol)
        assert bf * n <= max(2.0, max_bad_frac * n) + 1e-9, \
            f"{name}: {bf * n:.0f} of {n} elements differ from the reference by more than {rtol} (relative to max)"
    else:
        n_ours, n_theirs = bad_fraction(mine, truth64, rtol) * n, bad_fraction(ref32, truth64, rtol) * n
        assert n_ours <= 2 + 3 * n_theirs + 1e-9, f"{name}: {n_ours:.0f} elements off the exact trajectory, reference has {n_theirs:.0f}"
        ours, theirs = rel_to_max(mine, truth64), rel_to_max(ref32, truth64)
        assert ours <= rtol + 10.0 * theirs, f"{name}: ours {ours:.2e} vs exact, reference {theirs:.2e} vs exact"
