"""On-GPU batch assembly: HBM-resident feature stores + a row-gather kernel in place of the reference's CPU collate.

Reference (src/towers/pairs/unified_bid_data_loader.py): the Dataset yields {"notice_idx", "company_idx"} per pair (:461-504);
create_collate_fn (:630-684) fancy-indexes dataset.notice_store / company_store ('dense_projected' [N,D] float,
'categorical' [N,K] int) on the CPU, builds the KJTs with _build_batch_kjt (:827-841) and scripts/train.py:261-273 copies the
batch to the GPU. Here the stores live in HBM (`DeviceFeatureStore`), a batch is two `tt_batch_gather` launches, and only the
index vectors cross PCIe. The result is the same batch dict ({"notice": {"dense", "kjt"}, "company": {...}}, sample-major KJT
values, unit lengths) bit for bit. No CPU fallback: the stores must be CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import native
from .kjt import KeyedJaggedTensor
from .native import TT_BF16, TT_F32, TT_LAYOUT_KEY_MAJOR, TT_LAYOUT_SAMPLE_MAJOR, ptr

TT_F16, TT_I64, TT_I32 = 2, 3, 4
_DENSE_TT = {torch.float32: TT_F32, torch.bfloat16: TT_BF16, torch.float16: TT_F16}
_CAT_TT = {torch.int64: TT_I64, torch.int32: TT_I32}


def _as_tensor(x) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x


class DeviceFeatureStore:
    """One tower's feature store resident in HBM: the reference's dataset.notice_store / dataset.company_store
    ({'dense_projected': [N,D] float, 'categorical': [N,K] int}, unified_bid_data_loader.py:640-668).
    `dense_dtype` / `cat_dtype` choose the resident precision (fp32 + int64 reproduce the reference's batch bit for bit;
    bf16/f16 + int32 halve the footprint; the gather always emits f32 + int64)."""

    def __init__(self, dense_projected, categorical, keys: Sequence[str], device, dense_dtype: torch.dtype = torch.float32,
                 cat_dtype: torch.dtype = torch.int64):
        native.lib()  # fail loudly without the CUDA library
        dense, cat = _as_tensor(dense_projected), _as_tensor(categorical)
        if dense.dim() != 2 or cat.dim() != 2 or dense.shape[0] != cat.shape[0]:
            raise ValueError("dense_projected [N,D] and categorical [N,K] must have the same number of rows")
        if cat.shape[1] != len(keys):
            raise ValueError(f"categorical has {cat.shape[1]} columns for {len(keys)} keys")
        if dense_dtype not in _DENSE_TT or cat_dtype not in _CAT_TT:
            raise ValueError("dense_dtype must be float32/bfloat16/float16 and cat_dtype int64/int32")
        if cat_dtype == torch.int32 and cat.numel() and (int(cat.max()) >= 2 ** 31 or int(cat.min()) < -2 ** 31):
            raise ValueError("categorical ids do not fit int32")
        self.device = torch.device(device)
        self.keys = list(keys)
        self.n_rows, self.D, self.K = int(dense.shape[0]), int(dense.shape[1]), int(cat.shape[1])
        # dense rows padded to a multiple of 4 elements so every row starts 16-byte (8-byte for 16-bit types) aligned
        ld = (self.D + 3) // 4 * 4
        self.dense = torch.zeros(self.n_rows, ld, dtype=dense_dtype, device=self.device)
        self.dense[:, :self.D].copy_(dense.to(self.device, non_blocking=True))
        self.cat = cat.to(self.device).to(cat_dtype).contiguous()

    @classmethod
    def from_dataset_store(cls, store: dict, keys: Sequence[str], device, **kw) -> "DeviceFeatureStore":
        return cls(store["dense_projected"], store["categorical"], keys, device, **kw)

    def nbytes(self) -> int:
        return self.dense.numel() * self.dense.element_size() + self.cat.numel() * self.cat.element_size()

    def gather(self, idx: torch.Tensor, dense_out: Optional[torch.Tensor] = None, values_out: Optional[torch.Tensor] = None,
               layout: str = "sample_major", d_err: Optional[torch.Tensor] = None):
        """(dense f32 [B,D], values i64 [B*K]) of rows `idx` (int64 CUDA tensor) — numpy's store[idx] of the reference collate."""
        if idx.dtype != torch.int64 or idx.device != self.device:
            raise ValueError(f"idx must be an int64 tensor on {self.device}")
        B = idx.numel()
        if dense_out is None:
            dense_out = torch.empty(B, self.D, dtype=torch.float32, device=self.device)
        if values_out is None:
            values_out = torch.empty(B * self.K, dtype=torch.int64, device=self.device)
        if dense_out.dtype != torch.float32 or dense_out.shape != (B, self.D) or dense_out.stride(1) != 1:
            raise ValueError("dense_out must be f32 [B, D] with unit column stride")
        if values_out.dtype != torch.int64 or values_out.numel() != B * self.K or not values_out.is_contiguous():
            raise ValueError("values_out must be contiguous int64 [B*K]")
        if B == 0:
            return dense_out, values_out
        out_ld = dense_out.stride(0) if B > 1 else max(self.D, dense_out.stride(0))
        native.lib().call("tt_batch_gather", ptr(self.dense), _DENSE_TT[self.dense.dtype], self.dense.stride(0), self.D, ptr(self.cat),
                          _CAT_TT[self.cat.dtype], self.cat.stride(0), self.K, self.n_rows, ptr(idx), B, ptr(dense_out), out_ld,
                          ptr(values_out), TT_LAYOUT_KEY_MAJOR if layout == "key_major" else TT_LAYOUT_SAMPLE_MAJOR, ptr(d_err),
                          torch.cuda.current_stream(self.device).cuda_stream)
        return dense_out, values_out


class GpuBatchAssembler:
    """collate on the GPU: `assemble(notice_idx, company_idx)` returns the reference's batch dict; `assemble_into(gstep, ...)`
    writes straight into the static input buffers of a captured `GraphedFusedStep` (no intermediate batch at all)."""

    def __init__(self, notice_store: DeviceFeatureStore, company_store: DeviceFeatureStore, layout: str = "sample_major"):
        if notice_store.device != company_store.device:
            raise ValueError("both stores must live on the same device")
        self.notice, self.company, self.layout = notice_store, company_store, layout
        self.device = notice_store.device
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.launches = 0

    def _idx(self, idx) -> torch.Tensor:
        if not torch.is_tensor(idx):
            idx = torch.as_tensor(np.asarray(idx, dtype=np.int64))
        return idx.to(self.device, dtype=torch.int64, non_blocking=True)

    def check(self):
        """Raise like numpy fancy-indexing would if any index of an earlier assemble() was out of range (device sync)."""
        n = int(self.err.item())
        if n:
            self.err.zero_()
            raise IndexError(f"{n} feature-store indices out of range (unified_bid_data_loader.py:466 contract)")

    def assemble(self, notice_idx, company_idx, check: bool = False) -> Dict[str, Dict[str, object]]:
        ni, ci = self._idx(notice_idx), self._idx(company_idx)
        if ni.numel() != ci.numel():
            raise ValueError(f"Batch size mismatch: notice={ni.numel()}, company={ci.numel()}")
        nd, nv = self.notice.gather(ni, layout=self.layout, d_err=self.err)
        cd, cv = self.company.gather(ci, layout=self.layout, d_err=self.err)
        self.launches += 2
        if check:
            self.check()
        return {"notice": {"dense": nd, "kjt": KeyedJaggedTensor(self.notice.keys, nv, layout=self.layout)},
                "company": {"dense": cd, "kjt": KeyedJaggedTensor(self.company.keys, cv, layout=self.layout)}}

    def assemble_into(self, gstep, notice_idx, company_idx):
        """Gather straight into gstep.static = [notice dense, notice values, company dense, company values]."""
        ni, ci = self._idx(notice_idx), self._idx(company_idx)
        s = gstep.static
        if ni.numel() != s[0].shape[0] or ci.numel() != s[2].shape[0]:
            raise ValueError("the captured step has a fixed batch size")
        self.notice.gather(ni, s[0], s[1], layout=self.layout, d_err=self.err)
        self.company.gather(ci, s[2], s[3], layout=self.layout, d_err=self.err)
        self.launches += 2


def create_collate_fn(assembler: GpuBatchAssembler):
    """Drop-in for unified_bid_data_loader.create_collate_fn(dataset): takes the Dataset's items
    ([{"notice_idx": i, "company_idx": j, ...}, ...]) and returns the batch dict — already on the GPU."""

    def collate_fn_device(batch: List[Dict]) -> Dict[str, Dict[str, object]]:
        notice_indices = np.fromiter((item["notice_idx"] for item in batch), dtype=np.int64, count=len(batch))
        company_indices = np.fromiter((item["company_idx"] for item in batch), dtype=np.int64, count=len(batch))
        return assembler.assemble(torch.from_numpy(notice_indices), torch.from_numpy(company_indices))

    return collate_fn_device
