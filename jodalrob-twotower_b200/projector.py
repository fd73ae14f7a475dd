"""Feature pre-projection: the reference's frozen FeatureProjector
(src/torchrec_preprocess/feature_projector.py:4-28 — numeric 29 -> 128 -> 128 and text 768 -> 128 -> 128, Linear-ReLU-Linear each)
and the store projection loop around it (feature_preprocessor.py:150-233: project in batches, torch.cat([numeric, text...]),
copy to the CPU, np.concatenate). Same constructor, parameter names and initialisation; the four Linears run through the
library's fp32 GEMM (tt_linear_forward, bias and ReLU fused), and `project_store` writes every part straight into its column
range of ONE [N, 128 + 128*n_text] device buffer — which can be the dense store of loader.DeviceFeatureStore, so the projected
features never visit the host. No CPU fallback."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import native
from .native import TT_F32, ptr


class FeatureProjector(nn.Module):
    def __init__(self, num_dim: int, text_dim: int, num_proj_dim: int = 128, text_proj_dim: int = 128, device="cuda:0"):
        super().__init__()
        # identical module tree to the reference: same state_dict keys, same RNG consumption at construction
        self.num_proj = nn.Sequential(nn.Linear(num_dim, num_proj_dim), nn.ReLU(), nn.Linear(num_proj_dim, num_proj_dim))
        self.text_proj = nn.Sequential(nn.Linear(text_dim, text_proj_dim), nn.ReLU(), nn.Linear(text_proj_dim, text_proj_dim))
        self.to(device)
        self.requires_grad_(False)  # frozen in the reference's pipeline (used under no_grad only)
        native.lib()

    def _mlp(self, seq: nn.Sequential, x: torch.Tensor, out: torch.Tensor):
        """out[M, N2] (a possibly strided column range of a wider buffer) = Linear2(relu(Linear1(x)))"""
        lib = native.lib()
        if not x.is_cuda:
            raise ValueError("inputs must be CUDA tensors (there is no CPU fallback)")
        x = x.float().contiguous()
        l1, l2 = seq[0], seq[2]
        M = x.shape[0]
        if M == 0:
            return out
        s = torch.cuda.current_stream(x.device).cuda_stream
        hid = torch.empty(M, l1.out_features, dtype=torch.float32, device=x.device)
        lib.call("tt_linear_forward", ptr(x), x.stride(0), ptr(l1.weight), ptr(l1.bias), ptr(hid), hid.stride(0), M, l1.out_features,
                 l1.in_features, TT_F32, 1, s)
        lib.call("tt_linear_forward", ptr(hid), hid.stride(0), ptr(l2.weight), ptr(l2.bias), ptr(out), out.stride(0), M,
                 l2.out_features, l2.in_features, TT_F32, 0, s)
        return out

    def forward(self, dense: Optional[torch.Tensor], text_dict: Dict[str, torch.Tensor]):
        dense_proj = None
        if dense is not None:
            dense_proj = self._mlp(self.num_proj, dense, torch.empty(dense.shape[0], self.num_proj[2].out_features,
                                                                     dtype=torch.float32, device=dense.device))
        text_proj = {col: self._mlp(self.text_proj, x, torch.empty(x.shape[0], self.text_proj[2].out_features, dtype=torch.float32,
                                                                   device=x.device)) for col, x in text_dict.items()}
        return dense_proj, text_proj

    def project_store(self, numeric, text: Optional[Dict[str, object]] = None, text_cols: Optional[Sequence[str]] = None,
                      out: Optional[torch.Tensor] = None, batch_size: int = 65536) -> Optional[torch.Tensor]:
        """dense_projected [N, num_proj (+ text_proj per text column)] on the device (feature_preprocessor.py:150-233):
        numeric first, then the text columns in `text_cols` order. Inputs may be numpy arrays (staged per batch) or CUDA tensors."""
        text = text or {}
        cols = [c for c in (text_cols if text_cols else list(text.keys())) if c in text]
        dev = self.num_proj[0].weight.device
        n = len(numeric) if numeric is not None else (len(next(iter(text.values()))) if text else 0)
        wn, wt = self.num_proj[2].out_features, self.text_proj[2].out_features
        width = (wn if numeric is not None else 0) + wt * len(cols)
        if width == 0:
            return None
        if out is None:
            out = torch.empty(n, width, dtype=torch.float32, device=dev)
        if out.shape[0] != n or out.shape[1] < width or out.dtype != torch.float32 or out.stride(1) != 1:
            raise ValueError("out must be f32 [N, >= projected width] with unit column stride")
        stage = lambda a, lo, hi: (torch.from_numpy(np.ascontiguousarray(a[lo:hi])) if isinstance(a, np.ndarray) else a[lo:hi]).to(dev).float()
        with torch.no_grad():
            for lo in range(0, n, batch_size):
                hi = min(lo + batch_size, n)
                c0 = 0
                if numeric is not None:
                    self._mlp(self.num_proj, stage(numeric, lo, hi), out[lo:hi, c0:c0 + wn])
                    c0 += wn
                for col in cols:
                    self._mlp(self.text_proj, stage(text[col], lo, hi), out[lo:hi, c0:c0 + wt])
                    c0 += wt
        return out
