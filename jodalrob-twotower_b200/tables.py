"""EmbeddingBagCollection-style pooled tables driven straight through the C ABI.

The train step of the reference only ever has one id per bag (unified_bid_data_loader.py:827-841), and the engine's step
is specialised for that. BASELINE.json's large-table configurations (10 M rows per table; 500 M rows, E=64, multi-hot bags
with pooling factor 20) exercise the same three entry points at sizes where they are HBM-bound:

    tt_ebc_forward (jagged bags, sum pooling)  ->  tt_ebc_sort_rows  ->  tt_ebc_backward_adam (sparse | dense_exact)

`PooledTables` owns one arena [R, E] (+ Adam moments) for K tables and runs those calls; it is what the parity tests and
bench.py's `tables` section use. Oracle: torch embedding_bag(mode="sum") + the explicit Adam rule (SURVEY.md §8c, L>1 row).
There is no CPU fallback: construction loads the CUDA library.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import native
from .native import HYPER_BYTES, TT_ADAM_DENSE_EXACT, TT_ADAM_SPARSE, TT_BF16, TT_F32, TT_LAYOUT_KEY_MAJOR, TT_LAYOUT_SAMPLE_MAJOR, ptr


class PooledTables:
    def __init__(self, vocab: Sequence[int], emb_dim: int, device, optimizer_mode: str = "sparse", init_std: float = 1.0,
                 seed: int = 0, with_optimizer: bool = True):
        self.lib = native.lib()
        self.device = torch.device(device)
        self.vocab_host = [int(v) for v in vocab]
        self.K, self.E = len(self.vocab_host), int(emb_dim)
        if self.E % 4:
            raise ValueError("emb_dim must be a multiple of 4")
        base, r = [], 0
        for v in self.vocab_host:
            base.append(r)
            r += v
        self.total_rows = r
        if r >= 2 ** 31:
            raise ValueError("at most 2^31-1 arena rows per GPU shard")
        self.row_base_host = base
        self.row_base = torch.tensor(base, dtype=torch.int64, device=self.device)
        self.vocab = torch.tensor(self.vocab_host, dtype=torch.int32, device=self.device)
        g = torch.Generator(device=self.device).manual_seed(seed)
        self.w = torch.empty(r, self.E, dtype=torch.float32, device=self.device)
        if init_std > 0:
            self.w.normal_(0.0, init_std, generator=g)
        else:
            self.w.zero_()
        self.mode = {"dense_exact": TT_ADAM_DENSE_EXACT, "sparse": TT_ADAM_SPARSE}[optimizer_mode]
        self.m = torch.zeros_like(self.w) if with_optimizer else None
        self.v = torch.zeros_like(self.w) if with_optimizer else None
        self.row_begin = torch.zeros(r, dtype=torch.int32, device=self.device)
        self.row_end = torch.zeros(r, dtype=torch.int32, device=self.device)
        self.row_stamp = torch.zeros(r, dtype=torch.int32, device=self.device)
        self.stamp = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.hyper = torch.zeros(HYPER_BYTES, dtype=torch.uint8, device=self.device)
        self._ws_n = -1
        self.launches = 0

    def table(self, k: int, which: str = "w") -> torch.Tensor:
        b = self.row_base_host[k]
        return getattr(self, which)[b:b + self.vocab_host[k]]

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, n: int):
        if n == self._ws_n:
            return
        dev = self.device
        self.rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.slots = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.sorted_rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.sorted_slots = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.sort_ws = torch.empty(max(self.lib.query("tt_ebc_sort_workspace_bytes", n, self.total_rows), 16), dtype=torch.uint8, device=dev)
        self.piece_ws = torch.empty(max(self.lib.query("tt_ebc_piece_workspace_bytes", n, self.E), 16), dtype=torch.uint8, device=dev)
        self._ws_n = n

    # ------------------------------------------------------------------ forward
    def forward(self, values: torch.Tensor, B: int, bag_offsets: Optional[torch.Tensor] = None, layout: str = "sample_major",
                out: Optional[torch.Tensor] = None, out_dtype=torch.float32, record: bool = True) -> torch.Tensor:
        """pooled [B, K*E]; bag j (sample-major j = b*K+k) sums rows values[bag_offsets[j]:bag_offsets[j+1]] of table k,
        ids clamped to [0, V_k-1] (cat_embed.py:115-117). `record` keeps (row, slot) of every value for the backward."""
        if values.dtype != torch.int64:
            raise ValueError("values must be int64")
        if bag_offsets is not None and (bag_offsets.dtype != torch.int64 or bag_offsets.numel() != B * self.K + 1):
            raise ValueError("bag_offsets must be int64 [B*K+1]")
        n = values.numel()
        if bag_offsets is None and n != B * self.K:
            raise ValueError(f"without bag_offsets there must be exactly B*K values ({n} != {B * self.K})")
        if out is None:
            out = torch.empty(B, self.K * self.E, dtype=out_dtype, device=self.device)
        if record:
            self._workspace(n)
        self.lib.call("tt_ebc_forward", ptr(self.w), ptr(self.row_base), ptr(self.vocab), ptr(values), ptr(bag_offsets), n, B, self.K,
                      self.E, TT_LAYOUT_KEY_MAJOR if layout == "key_major" else TT_LAYOUT_SAMPLE_MAJOR, ptr(out),
                      TT_F32 if out.dtype == torch.float32 else TT_BF16, out.stride(0), 0, ptr(self.rows) if record else 0,
                      ptr(self.slots) if record else 0, 0, self._stream())
        self.launches += 1
        self._n, self._B = n, B
        return out

    # ------------------------------------------------------------------ backward + optimizer
    def hyper_advance(self, lr=1e-3, warmup_steps=0, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        self.lib.call("tt_hyper_advance", ptr(self.hyper), lr, warmup_steps, betas[0], betas[1], eps, weight_decay, 0, self._stream())
        self.launches += 1

    def sort(self):
        """deterministic sort-by-row dedup of the (row, slot) pairs recorded by the last forward"""
        self.stamp.add_(1)
        self.lib.call("tt_ebc_sort_rows", ptr(self.rows), ptr(self.slots), self._n, self.total_rows, ptr(self.sort_ws), self.sort_ws.numel(),
                      ptr(self.sorted_rows), ptr(self.sorted_slots), ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp), 0,
                      ptr(self.stamp), 0, self._stream())
        bits = max(1, (self.total_rows - 1).bit_length())
        self.launches += 3 * ((bits + 7) // 8) + 1

    def backward_adam(self, grad: torch.Tensor, dense_grad_out: Optional[torch.Tensor] = None, apply_adam: bool = True):
        """grad: pooled-output gradient [B, K*E] (f32 or bf16). Segment-sum per row in fixed order, fused Adam row update
        (call hyper_advance first); `dense_grad_out` [R, E] additionally/only materialises the dense gradient."""
        B = self._B
        if grad.shape != (B, self.K * self.E) or grad.stride(1) != 1:
            raise ValueError("grad must be [B, K*E] with unit column stride")
        if apply_adam and self.m is None:
            raise ValueError("tables were built without optimizer state")
        self.lib.call("tt_ebc_backward_adam", ptr(self.w) if apply_adam else 0, ptr(self.m) if apply_adam else 0,
                      ptr(self.v) if apply_adam else 0, ptr(dense_grad_out), self.total_rows, self.E, ptr(self.sorted_rows),
                      ptr(self.sorted_slots), self._n, ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp), 0, ptr(self.stamp),
                      ptr(grad), grad.stride(0), 0, self.K, B * self.K, 0, 0, 0, 1, 0,
                      TT_F32 if grad.dtype == torch.float32 else TT_BF16, ptr(self.piece_ws), self.piece_ws.numel(), ptr(self.hyper),
                      self.mode if apply_adam else TT_ADAM_DENSE_EXACT, 0, 0, 0, 0, 0, self._stream())
        self.launches += 3

    # ------------------------------------------------------------------ algorithmic bytes
    def forward_bytes(self, B: int, n_values: int, out_elem_bytes: int = 4, jagged: bool = True) -> int:
        return n_values * (self.E * 4 + 8) + B * self.K * self.E * out_elem_bytes + ((B * self.K + 1) * 8 if jagged else 0)

    def update_bytes(self, B: int, n_values: int, unique_rows: int, grad_elem_bytes: int = 4) -> int:
        rows = self.total_rows if self.mode == TT_ADAM_DENSE_EXACT else unique_rows
        return B * self.K * self.E * grad_elem_bytes + n_values * 8 + rows * self.E * 24
