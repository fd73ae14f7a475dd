"""Thin functional wrappers over single C-ABI entry points (used by the façade and by the kernel tests)."""
from __future__ import annotations

from typing import Optional

import torch

from . import native
from .native import TT_BF16, TT_F32, TT_LAYOUT_KEY_MAJOR, TT_LAYOUT_SAMPLE_MAJOR, ptr


def _lay(layout) -> int:
    if isinstance(layout, int):
        return layout
    return TT_LAYOUT_KEY_MAJOR if layout == "key_major" else TT_LAYOUT_SAMPLE_MAJOR


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def ebc_forward(engine, tower: int, values: torch.Tensor, B: int, layout="sample_major",
                bag_offsets: Optional[torch.Tensor] = None, out_dtype=torch.float32, want_rows: bool = False):
    """Pooled lookup of one tower's keys -> [B, K*E]. With `bag_offsets` (i64 [B*K+1]) bags are jagged and sum-pooled."""
    K = engine.K[tower]
    E = engine.cfg.emb_dim
    k0 = 0 if tower == 0 else engine.K[0]
    out = torch.empty(B, K * E, dtype=out_dtype, device=engine.device)
    n_values = values.numel()
    rows = torch.empty(max(n_values, 1), dtype=torch.int32, device=engine.device) if want_rows else None
    slots = torch.empty(max(n_values, 1), dtype=torch.int32, device=engine.device) if want_rows else None
    engine._call("tt_ebc_forward", ptr(engine.flat_w), engine.row_base.data_ptr() + 8 * k0, engine.vocab.data_ptr() + 4 * k0,
                 ptr(values), ptr(bag_offsets), n_values, B, K, E, _lay(layout), ptr(out),
                 TT_F32 if out_dtype == torch.float32 else TT_BF16, K * E, 0, ptr(rows), ptr(slots), 0, _stream(engine.device))
    return (out, rows, slots) if want_rows else out


def similarity_matrix(engine, n: torch.Tensor, c: torch.Tensor, inv_temp: float = 1.0) -> torch.Tensor:
    n = n.contiguous().float()
    c = c.contiguous().float()
    S = torch.empty(n.shape[0], c.shape[0], dtype=torch.float32, device=n.device)
    engine._call("tt_similarity_matrix", ptr(n), ptr(c), n.shape[0], c.shape[0], n.shape[1], TT_F32, float(inv_temp), ptr(S),
                 c.shape[0], _stream(n.device))
    return S
