"""Kernel orchestration for one two-tower train step on one GPU (host side of the C ABI).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic stage is an
`extern "C"` call into csrc/libtwotower_b200.so (include/twotower_b200.h).  There is no CPU path.

Data layout in HBM
  arena        f32 [R, E]   all embedding tables (notice keys then company keys) back to back; the
                            per-key nn.Embedding weights of the reference's state_dict are row views
  arena_m/v    f32 [R, E]   Adam moments, same layout
  dense_w/g/m/v f32 [P]     all Linear / BatchNorm tensors of both towers in one flat buffer
                            (one Adam launch; one all-reduce on multi-GPU)
  per tower    c  [B, H0+K*E] concat buffer (dense projection writes cols [0,H0), the EBC lookup the rest)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native
from .native import TT_ADAM_DENSE_EXACT, TT_ADAM_SPARSE, TT_F32, TT_LAYOUT_KEY_MAJOR, TT_LAYOUT_SAMPLE_MAJOR, ptr

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


@dataclass
class TwoTowerConfig:
    notice_keys: List[str]
    company_keys: List[str]
    notice_vocab: List[int]
    company_vocab: List[int]
    emb_dim: int = 64
    notice_dense_dim: int = 256
    company_dense_dim: int = 128
    hidden: List[int] = field(default_factory=lambda: [256, 128])
    final_dim: int = 128
    dropout: float = 0.2
    temperature: float = 1.0
    label_smoothing: float = 0.0

    def tower(self, t: int):
        return (("notice", self.notice_keys, self.notice_vocab, self.notice_dense_dim),
                ("company", self.company_keys, self.company_vocab, self.company_dense_dim))[t]


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class DenseLayout:
    """Names/shapes/offsets of every dense (non-embedding) parameter inside the flat buffer,
    in the reference's registration order (base_tower.py:59,87-97)."""

    def __init__(self, cfg: TwoTowerConfig):
        self.entries: List[Tuple[str, Tuple[int, ...], int]] = []
        off = 0
        E = cfg.emb_dim
        for t in range(2):
            name, keys, _, dense_dim = cfg.tower(t)
            pre = f"{name}_tower."
            shapes = [(pre + "dense_projection.weight", (cfg.hidden[0], dense_dim)),
                      (pre + "dense_projection.bias", (cfg.hidden[0],))]
            width = cfg.hidden[0] + len(keys) * E
            idx = 0
            for h in cfg.hidden[1:]:
                shapes += [(pre + f"mlp.{idx}.weight", (h, width)), (pre + f"mlp.{idx}.bias", (h,)),
                           (pre + f"mlp.{idx + 2}.weight", (h,)), (pre + f"mlp.{idx + 2}.bias", (h,))]
                width = h
                idx += 4
            shapes += [(pre + f"mlp.{idx}.weight", (cfg.final_dim, width)), (pre + f"mlp.{idx}.bias", (cfg.final_dim,))]
            for n, shp in shapes:
                self.entries.append((n, shp, off))
                off += _align(math.prod(shp))
        self.total = off
        self.index = {n: (shp, o) for n, shp, o in self.entries}

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        shp, off = self.index[name]
        return flat[off:off + math.prod(shp)].view(shp)


class _Ws:
    """Per-batch-size activation / scratch buffers."""
    pass


class TwoTowerEngine:
    def __init__(self, cfg: TwoTowerConfig, device, optimizer_mode: str = "dense_exact", value_layout: str = "sample_major",
                 dtype: str = "fp32"):
        self.cfg = cfg
        if dtype not in ("fp32", "bf16"):
            raise ValueError(f"dtype must be 'fp32' or 'bf16', got {dtype}")
        # bf16: tower outputs are stored in bf16 and the B x B logits / loss / gradients run on tcgen05 tensor cores
        # (fp32 accumulation in TMEM); master weights, Adam state and reductions stay fp32.
        self.dtype = dtype
        if dtype == "bf16" and cfg.final_dim not in (64, 128):
            raise ValueError("dtype='bf16' (tcgen05 sim-CE) needs final_embedding_dim 64 or 128")
        self.emb_tt = native.TT_BF16 if dtype == "bf16" else TT_F32
        self.emb_torch = torch.bfloat16 if dtype == "bf16" else torch.float32
        if dtype == "bf16":
            widths = [cfg.notice_dense_dim, cfg.company_dense_dim, cfg.final_dim, cfg.emb_dim] + list(cfg.hidden)
            if any(w % 8 for w in widths):
                raise ValueError("dtype='bf16' (TMA-fed tcgen05 GEMMs) needs every layer width to be a multiple of 8")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise native.TwoTowerNativeError("TwoTowerEngine needs a CUDA device: the hot path is sm_100a CUDA only, no CPU fallback")
        self.lib = native.lib()
        self.act = native.TT_BF16 if dtype == "bf16" else TT_F32
        self.act_torch = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.adam_mode = {"dense_exact": TT_ADAM_DENSE_EXACT, "sparse": TT_ADAM_SPARSE}[optimizer_mode]
        self.default_layout = value_layout
        E = cfg.emb_dim
        if E % 4 != 0 or cfg.hidden[0] % 4 != 0:
            raise ValueError("emb_dim and tower_hidden_dims[0] must be multiples of 4 (128-bit vector path)")
        dev = self.device
        vocab = list(cfg.notice_vocab) + list(cfg.company_vocab)
        base, acc = [], 0
        for v in vocab:
            base.append(acc)
            acc += int(v)
        self.total_rows = acc
        self.K = (len(cfg.notice_keys), len(cfg.company_keys))
        self.row_base_host = base
        self.row_base = torch.tensor(base, dtype=torch.int64, device=dev)
        self.vocab = torch.tensor(vocab, dtype=torch.int32, device=dev)
        self.arena = torch.empty(self.total_rows, E, dtype=torch.float32, device=dev)
        self.arena_m = torch.zeros_like(self.arena)
        self.arena_v = torch.zeros_like(self.arena)
        self.arena_grad: Optional[torch.Tensor] = None
        self.layout = DenseLayout(cfg)
        P = self.layout.total
        self.dense_w = torch.zeros(P, dtype=torch.float32, device=dev)
        self.dense_g = torch.zeros(P, dtype=torch.float32, device=dev)
        self.dense_m = torch.zeros(P, dtype=torch.float32, device=dev)
        self.dense_v = torch.zeros(P, dtype=torch.float32, device=dev)
        self.dense_w16 = torch.zeros(P, dtype=torch.bfloat16, device=dev) if dtype == "bf16" else None  # bf16 mirror for the GEMMs
        nblk = len(cfg.hidden) - 1
        self.bn_rm = [[torch.zeros(h, device=dev) for h in cfg.hidden[1:]] for _ in range(2)]
        self.bn_rv = [[torch.ones(h, device=dev) for h in cfg.hidden[1:]] for _ in range(2)]
        self.bn_nbt = [[torch.zeros((), dtype=torch.int64, device=dev) for _ in range(nblk)] for _ in range(2)]
        self.hyper = torch.zeros(native.HYPER_BYTES, dtype=torch.uint8, device=dev)
        self.stamp = torch.zeros(1, dtype=torch.int64, device=dev)
        self.row_begin = torch.zeros(self.total_rows, dtype=torch.int32, device=dev)
        self.row_end = torch.zeros(self.total_rows, dtype=torch.int32, device=dev)
        self.row_stamp = torch.zeros(self.total_rows, dtype=torch.int32, device=dev)
        self._ws: Dict[int, _Ws] = {}
        self.rng_seed = 0x5EED
        self.rng_stream_base = 0  # distributed engines offset it by rank so ranks draw different dropout masks
        # the two towers are independent until the loss: run the company tower on a side stream (fork/join with events;
        # inside a CUDA graph capture this becomes two parallel branches)
        self.two_streams = True
        self._side = torch.cuda.Stream(self.device)
        self.launches = 0  # kernels launched through the C ABI (bench.py's gpu_launches)
        self.profile: Optional[list] = None  # set to [] to collect (tag, n_kernels, ev0, ev1) per C-ABI call
        self.profile_tag = ""

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, name: str, *args, n_kernels: int = 1):
        self.launches += n_kernels
        prof = self.profile
        if prof is None:
            self.lib.call(name, *args)
            return
        # per-entry-point device timing (bench.py's kernel table): CUDA events on the launching stream
        st = torch.cuda.current_stream(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        self.lib.call(name, *args)
        e1.record(st)
        prof.append((f"{name}[{self.profile_tag}]" if self.profile_tag else name, n_kernels, e0, e1))

    def dense(self, name: str, which: str = "w") -> torch.Tensor:
        return self.layout.view({"w": self.dense_w, "g": self.dense_g, "m": self.dense_m, "v": self.dense_v}[which], name)

    def table_rows(self, t: int, k: int) -> Tuple[int, int]:
        gk = k if t == 0 else self.K[0] + k
        v = (self.cfg.notice_vocab if t == 0 else self.cfg.company_vocab)[k]
        return self.row_base_host[gk], self.row_base_host[gk] + v

    def workspace(self, B: int) -> _Ws:
        ws = self._ws.get(B)
        if ws is not None:
            return ws
        cfg, dev, E = self.cfg, self.device, self.cfg.emb_dim
        q = self.lib.query
        ws = _Ws()
        ws.B = B
        f32 = dict(dtype=torch.float32, device=dev)
        ws.c, ws.dc, ws.h, ws.y, ws.dh, ws.dy = [], [], [], [], [], []
        ws.mean, ws.rstd, ws.z, ws.e, ws.nrm, ws.de, ws.dz = [], [], [], [], [], [], []
        n_tot = B * (self.K[0] + self.K[1])
        ws.rows = torch.zeros(max(n_tot, 1), dtype=torch.int32, device=dev)
        ws.sorted_rows = torch.zeros_like(ws.rows)
        ws.sorted_slots = torch.zeros_like(ws.rows)
        max_bn = 0
        max_wg = 0
        for t in range(2):
            _, keys, _, dense_dim = cfg.tower(t)
            ld = cfg.hidden[0] + len(keys) * E
            act = dict(dtype=self.act_torch, device=dev)
            ws.c.append(torch.empty(B, ld, **act))
            ws.dc.append(torch.empty(B, ld, **act))
            ws.h.append([torch.empty(B, h, **act) for h in cfg.hidden[1:]])
            ws.y.append([torch.empty(B, h, **act) for h in cfg.hidden[1:]])
            ws.dh.append([torch.empty(B, h, **act) for h in cfg.hidden[1:]])
            ws.dy.append([torch.empty(B, h, **act) for h in cfg.hidden[1:]])
            if self.act != TT_F32:
                if t == 0:
                    ws.x16, ws.dz16 = [], []
                ws.x16.append(torch.empty(B, dense_dim, **act))
                ws.dz16.append(torch.empty(B, cfg.final_dim, **act))
            ws.mean.append([torch.empty(h, **f32) for h in cfg.hidden[1:]])
            ws.rstd.append([torch.empty(h, **f32) for h in cfg.hidden[1:]])
            ws.z.append(torch.empty(B, cfg.final_dim, **f32))
            ws.e.append(torch.empty(B, cfg.final_dim, dtype=self.emb_torch, device=dev))
            ws.nrm.append(torch.empty(B, **f32))
            ws.de.append(torch.empty(B, cfg.final_dim, **f32))
            ws.dz.append(torch.empty(B, cfg.final_dim, **f32))
            width = ld
            wgq = "tt_linear_wgrad_workspace_bytes" if self.act == TT_F32 else "tt_linear_wgrad_bf16_workspace_bytes"
            max_wg = max(max_wg, q(wgq, B, cfg.hidden[0], dense_dim))
            for h in cfg.hidden[1:]:
                max_bn = max(max_bn, q("tt_bn_workspace_bytes", B, h))
                max_wg = max(max_wg, q(wgq, B, h, width))
                width = h
            max_wg = max(max_wg, q(wgq, B, cfg.final_dim, width))
        # one scratch per tower so the two towers may run on different streams
        ws.bn_ws = [torch.empty(max(max_bn, 256), dtype=torch.uint8, device=dev) for _ in range(2)]
        ws.bn_sums = [torch.zeros(2 * max(cfg.hidden[1:] + [1]), dtype=torch.float64, device=dev) for _ in range(2)]
        ws.wg_ws = [torch.empty(max_wg, dtype=torch.uint8, device=dev) for _ in range(2)]
        ws.sim_ws = torch.empty(q("tt_simce_workspace_bytes", B, B, cfg.final_dim), dtype=torch.uint8, device=dev)
        ws.row_sumexp = torch.empty(B, **f32)
        ws.col_sumexp = torch.empty(B, **f32)
        ws.row_sum_s = torch.empty(B, **f32)
        ws.col_sum_s = torch.empty(B, **f32)
        ws.diag = torch.empty(B, **f32)
        ws.row_argmax = torch.empty(B, dtype=torch.int32, device=dev)
        ws.inv_row = torch.empty(B, **f32)
        ws.inv_col = torch.empty(B, **f32)
        ws.partial = torch.zeros(8, dtype=torch.float64, device=dev)
        ws.out5 = torch.zeros(5, **f32)
        ws.sort_ws = torch.empty(q("tt_ebc_sort_workspace_bytes", n_tot, self.total_rows), dtype=torch.uint8, device=dev)
        ws.piece_ws = torch.empty(q("tt_ebc_piece_workspace_bytes", n_tot, E), dtype=torch.uint8, device=dev)
        self._ws[B] = ws
        return ws

    # ------------------------------------------------------------------ Linear dispatch (fp32 SIMT / bf16 tcgen05)
    def _w16(self, name: str) -> torch.Tensor:
        return self.layout.view(self.dense_w16, name)

    def _lin_fwd(self, x, ldx, wname, bname, out, ldo, out_f32: bool, M, N, K, relu):
        s = self._stream()
        if self.act == TT_F32:
            self._call("tt_linear_forward", ptr(x), ldx, ptr(self.dense(wname)), ptr(self.dense(bname)), ptr(out), ldo, M, N, K,
                       TT_F32, relu, s)
        else:
            self._call("tt_linear_forward_bf16", ptr(x), ldx, ptr(self._w16(wname)), ptr(self.dense(bname)), ptr(out), ldo,
                       TT_F32 if out_f32 else native.TT_BF16, M, N, K, relu, s)

    def _lin_dgrad(self, dy, ldy, wname, dx, ldx, M, N, K):
        s = self._stream()
        if self.act == TT_F32:
            self._call("tt_linear_dgrad", ptr(dy), ldy, ptr(self.dense(wname)), ptr(dx), ldx, M, N, K, TT_F32, s)
        else:
            self._call("tt_linear_dgrad_bf16", ptr(dy), ldy, ptr(self._w16(wname)), ptr(dx), ldx, native.TT_BF16, M, N, K, s)

    def _lin_wgrad(self, dy, ldy, x, ldx, wname, bname, M, N, K, wsbuf):
        s = self._stream()
        if self.act == TT_F32:
            self._call("tt_linear_wgrad", ptr(dy), ldy, ptr(x), ldx, ptr(self.dense(wname, "g")), ptr(self.dense(bname, "g")), M, N, K,
                       TT_F32, ptr(wsbuf), wsbuf.numel(), s, n_kernels=4)
        else:
            self._call("tt_linear_wgrad_bf16", ptr(dy), ldy, ptr(x), ldx, ptr(self.dense(wname, "g")), ptr(self.dense(bname, "g")),
                       M, N, K, ptr(wsbuf), wsbuf.numel(), s, n_kernels=4)

    def refresh_bf16_weights(self):
        if self.dense_w16 is not None:
            self._call("tt_cast_bf16", ptr(self.dense_w), ptr(self.dense_w16), self.layout.total, self._stream())

    # ------------------------------------------------------------------ forward
    def _check_inputs(self, t: int, dense: torch.Tensor, values: torch.Tensor, B: int):
        name, keys, _, dense_dim = self.cfg.tower(t)
        if dense.dim() != 2 or dense.shape[1] != dense_dim:
            raise ValueError(f"{name} dense must be [B,{dense_dim}], got {tuple(dense.shape)}")
        if values.dtype != torch.int64:
            raise ValueError(f"{name} kjt values must be int64, got {values.dtype}")
        if dense.device != self.device or values.device != self.device:
            raise ValueError(f"{name} inputs must live on {self.device}")

    def embedding_forward(self, ws: _Ws, values: Sequence[Optional[torch.Tensor]], layout: int):
        """Pooled lookup of every key of the given towers (None = skip) straight into the concat buffers; also records
        the clamped arena rows for the backward. One id per bag, like the reference (unified_bid_data_loader.py:835);
        jagged bags go through ops.ebc_forward."""
        E, B, s = self.cfg.emb_dim, ws.B, self._stream()
        H0 = self.cfg.hidden[0]
        for t, v in enumerate(values):
            if v is None:
                continue
            K = self.K[t]
            k0 = 0 if t == 0 else self.K[0]
            rows_off = 0 if t == 0 else B * self.K[0]
            self.profile_tag = self.cfg.tower(t)[0]
            self._call("tt_ebc_forward", ptr(self.arena), self.row_base.data_ptr() + 8 * k0, self.vocab.data_ptr() + 4 * k0,
                       ptr(v), 0, B * K, B, K, E, layout, ptr(ws.c[t]), self.act, H0 + K * E, H0,
                       ws.rows.data_ptr() + 4 * rows_off, 0, 0, s)

    def tower_forward(self, t: int, ws: _Ws, dense: torch.Tensor, values: Optional[torch.Tensor], train: bool,
                      layout: int = TT_LAYOUT_SAMPLE_MAJOR, masks: Optional[Sequence[torch.Tensor]] = None):
        """Dense projection + MLP + L2 norm of one tower. If `values` is given the embedding lookup of this tower is
        done here first (stand-alone tower call); the train step looks both towers up front via embedding_forward."""
        if values is not None:
            self.embedding_forward(ws, [values if i == t else None for i in range(2)], layout)
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        name, keys, _, dense_dim = cfg.tower(t)
        self.profile_tag = name
        K = len(keys)
        k0 = 0 if t == 0 else self.K[0]
        H0 = cfg.hidden[0]
        ld = H0 + K * E
        pre = f"{name}_tower."
        dense = dense.contiguous()
        if dense.dtype != torch.float32:
            dense = dense.float()
        if self.act != TT_F32:
            self._call("tt_cast_bf16", ptr(dense), ptr(ws.x16[t]), B * dense_dim, s)
            dense = ws.x16[t]
        self._lin_fwd(dense, dense_dim, pre + "dense_projection.weight", pre + "dense_projection.bias", ws.c[t], ld, False, B, H0,
                      dense_dim, 0)
        x, width, idx = ws.c[t], ld, 0
        for bi, h in enumerate(cfg.hidden[1:]):
            self._lin_fwd(x, width, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", ws.h[t][bi], h, False, B, h, width, 1)
            gamma, beta = self.dense(pre + f"mlp.{idx + 2}.weight"), self.dense(pre + f"mlp.{idx + 2}.bias")
            if train:
                self._call("tt_bn_stats", ptr(ws.h[t][bi]), h, B, h, self.act, ptr(ws.bn_sums[t]), ptr(ws.bn_ws[t]),
                           ws.bn_ws[t].numel(), s, n_kernels=2)
                count = self.bn_allreduce(ws.bn_sums[t], 2 * h, B)
                self._call("tt_bn_finalize", ptr(ws.bn_sums[t]), count, h, BN_EPS, BN_MOMENTUM, ptr(ws.mean[t][bi]),
                           ptr(ws.rstd[t][bi]), ptr(self.bn_rm[t][bi]), ptr(self.bn_rv[t][bi]), ptr(self.bn_nbt[t][bi]), s)
                mean = ws.mean[t][bi]
                p = cfg.dropout
            else:
                self._call("tt_bn_rstd_from_var", ptr(self.bn_rv[t][bi]), h, BN_EPS, ptr(ws.rstd[t][bi]), s)
                mean = self.bn_rm[t][bi]
                p = 0.0
            mask = masks[bi] if (masks is not None and p > 0) else None
            self._call("tt_bn_apply", ptr(ws.h[t][bi]), h, B, h, self.act, ptr(mean), ptr(ws.rstd[t][bi]), ptr(gamma),
                       ptr(beta), p, ptr(mask), ptr(self.hyper), self.rng_stream_base + t * 16 + bi, ptr(ws.y[t][bi]), h, s)
            x, width, idx = ws.y[t][bi], h, idx + 4
        F = cfg.final_dim
        self._lin_fwd(x, width, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", ws.z[t], F, True, B, F, width, 0)
        self._call("tt_l2norm_forward", ptr(ws.z[t]), B, F, ptr(ws.e[t]), self.emb_tt, ptr(ws.nrm[t]), s)
        return ws.e[t]

    def bn_allreduce(self, sums: torch.Tensor, n: int, local_count: int) -> int:
        """Single GPU: nothing to do. Distributed engines all-reduce `sums[:n]` and return the global count."""
        return local_count

    def loss_forward(self, ws: _Ws):
        cfg, B, s = self.cfg, ws.B, self._stream()
        self.profile_tag = ""
        inv_t = 1.0 / cfg.temperature
        self._call("tt_simce_forward", ptr(ws.e[0]), ptr(ws.e[1]), B, B, cfg.final_dim, self.emb_tt, inv_t, 0, ptr(ws.row_sumexp),
                   ptr(ws.col_sumexp), ptr(ws.row_sum_s), ptr(ws.col_sum_s), ptr(ws.diag), ptr(ws.row_argmax),
                   ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=3)
        self._call("tt_simce_finalize", ptr(ws.row_sumexp), ptr(ws.col_sumexp), ptr(ws.row_sum_s), ptr(ws.col_sum_s),
                   ptr(ws.diag), ptr(ws.row_argmax), B, B, 0, B, inv_t, cfg.label_smoothing, ptr(ws.inv_row),
                   ptr(ws.inv_col), ptr(ws.partial), s)
        self._call("tt_simce_scalars", ptr(ws.partial), B, ptr(ws.out5), s)
        return ws.out5

    def forward(self, notice_dense, notice_values, company_dense, company_values, train: bool = True,
                layout: Optional[str] = None, masks=None) -> Tuple[torch.Tensor, _Ws]:
        B = notice_dense.shape[0]
        if company_dense.shape[0] != B:
            raise ValueError(f"notice/company batch sizes differ: {B} vs {company_dense.shape[0]}")
        self._check_inputs(0, notice_dense, notice_values, B)
        self._check_inputs(1, company_dense, company_values, B)
        for t, v in ((0, notice_values), (1, company_values)):
            if v.numel() != B * self.K[t]:
                raise ValueError(f"kjt values hold {v.numel()} ids, expected B*K = {B * self.K[t]}")
        lay = TT_LAYOUT_KEY_MAJOR if (layout or self.default_layout) == "key_major" else TT_LAYOUT_SAMPLE_MAJOR
        ws = self.workspace(B)
        ws.inputs = (notice_dense.contiguous(), company_dense.contiguous())
        self.refresh_bf16_weights()
        self.embedding_forward(ws, [notice_values, company_values], lay)
        m0, m1 = (None, None) if masks is None else (masks[0], masks[1])
        if self.two_streams:
            main = torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.tower_forward(1, ws, company_dense, None, train, lay, masks=m1)
            self.tower_forward(0, ws, notice_dense, None, train, lay, masks=m0)
            main.wait_stream(self._side)
        else:
            self.tower_forward(0, ws, notice_dense, None, train, lay, masks=m0)
            self.tower_forward(1, ws, company_dense, None, train, lay, masks=m1)
        out5 = self.loss_forward(ws)
        ws.train = train
        ws.masks = masks
        return out5, ws

    # ------------------------------------------------------------------ backward
    def loss_backward(self, ws: _Ws, dloss: float = 1.0):
        cfg, B, s = self.cfg, ws.B, self._stream()
        self.profile_tag = ""
        inv_t = 1.0 / cfg.temperature
        gscale = dloss * 0.5 / B * inv_t
        F = cfg.final_dim
        self._call("tt_simce_backward", ptr(ws.e[0]), ptr(ws.e[1]), B, B, F, self.emb_tt, inv_t, 0, ptr(ws.inv_row), ptr(ws.inv_col),
                   gscale, cfg.label_smoothing, B, ptr(ws.de[0]), ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=2)
        self._call("tt_simce_backward", ptr(ws.e[1]), ptr(ws.e[0]), B, B, F, self.emb_tt, inv_t, 0, ptr(ws.inv_col), ptr(ws.inv_row),
                   gscale, cfg.label_smoothing, B, ptr(ws.de[1]), ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=2)

    def tower_backward(self, t: int, ws: _Ws):
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        name, keys, _, dense_dim = cfg.tower(t)
        self.profile_tag = name + ".bwd"
        H0, F = cfg.hidden[0], cfg.final_dim
        ld = H0 + len(keys) * E
        pre = f"{name}_tower."
        train = ws.train
        nblk = len(cfg.hidden) - 1
        widths = [ld] + list(cfg.hidden[1:])
        self._call("tt_l2norm_backward", ptr(ws.de[t]), ptr(ws.e[t]), self.emb_tt, ptr(ws.nrm[t]), B, F, ptr(ws.dz[t]), s)
        idx = 4 * nblk
        xlast = ws.y[t][nblk - 1] if nblk > 0 else ws.c[t]
        wl = widths[-1]
        dz = ws.dz[t]
        if self.act != TT_F32:
            self._call("tt_cast_bf16", ptr(ws.dz[t]), ptr(ws.dz16[t]), B * F, s)
            dz = ws.dz16[t]
        self._lin_wgrad(dz, F, xlast, wl, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", B, F, wl, ws.wg_ws[t])
        dx = ws.dy[t][nblk - 1] if nblk > 0 else ws.dc[t]
        self._lin_dgrad(dz, F, pre + f"mlp.{idx}.weight", dx, wl, B, F, wl)
        for bi in range(nblk - 1, -1, -1):
            h = cfg.hidden[1 + bi]
            idx = 4 * bi
            p = cfg.dropout if train else 0.0
            mask = ws.masks[t][bi] if (ws.masks is not None and p > 0) else None
            mean = ws.mean[t][bi] if train else self.bn_rm[t][bi]
            gamma = self.dense(pre + f"mlp.{idx + 2}.weight")
            self._call("tt_bn_backward_reduce", ptr(ws.dy[t][bi]), h, ptr(ws.h[t][bi]), h, B, h, self.act, ptr(mean),
                       ptr(ws.rstd[t][bi]), p, ptr(mask), ptr(self.hyper), self.rng_stream_base + t * 16 + bi, ptr(ws.bn_sums[t]),
                       ptr(ws.bn_ws[t]), ws.bn_ws[t].numel(), s, n_kernels=2)
            count = self.bn_allreduce(ws.bn_sums[t], 2 * h, B)
            self._call("tt_bn_backward_apply", ptr(ws.dy[t][bi]), h, ptr(ws.h[t][bi]), h, B, h, self.act, ptr(mean),
                       ptr(ws.rstd[t][bi]), ptr(gamma), p, ptr(mask), ptr(self.hyper), self.rng_stream_base + t * 16 + bi, ptr(ws.bn_sums[t]),
                       count, 1 if train else 0, ptr(self.dense(pre + f"mlp.{idx + 2}.weight", "g")),
                       ptr(self.dense(pre + f"mlp.{idx + 2}.bias", "g")), ptr(ws.dh[t][bi]), h, s)
            xin = ws.y[t][bi - 1] if bi > 0 else ws.c[t]
            win = widths[bi]
            self._lin_wgrad(ws.dh[t][bi], h, xin, win, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", B, h, win, ws.wg_ws[t])
            dxin = ws.dy[t][bi - 1] if bi > 0 else ws.dc[t]
            self._lin_dgrad(ws.dh[t][bi], h, pre + f"mlp.{idx}.weight", dxin, win, B, h, win)
        if self.act != TT_F32:
            dense_in = ws.x16[t]
        else:
            dense_in = ws.inputs[t]
            if dense_in.dtype != torch.float32:
                dense_in = dense_in.float()
        self._lin_wgrad(ws.dc[t], ld, dense_in, dense_dim, pre + "dense_projection.weight", pre + "dense_projection.bias", B, H0,
                        dense_dim, ws.wg_ws[t])

    def embedding_backward(self, ws: _Ws, apply_adam: bool, want_dense_grad: bool):
        """sort-by-row dedup of both towers' (row, slot) pairs, then segment-sum (+ fused Adam)."""
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        n0, n1 = B * self.K[0], B * self.K[1]
        n = n0 + n1
        H0 = cfg.hidden[0]
        self.profile_tag = ""
        self.stamp.add_(1)
        self._call("tt_ebc_sort_rows", ptr(ws.rows), 0, n, self.total_rows, ptr(ws.sort_ws), ws.sort_ws.numel(),
                   ptr(ws.sorted_rows), ptr(ws.sorted_slots), ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp),
                   0, ptr(self.stamp), 0, s, n_kernels=7)
        if want_dense_grad and self.arena_grad is None:
            self.arena_grad = torch.zeros_like(self.arena)
        self._call("tt_ebc_backward_adam", ptr(self.arena) if apply_adam else 0, ptr(self.arena_m) if apply_adam else 0,
                   ptr(self.arena_v) if apply_adam else 0, ptr(self.arena_grad) if want_dense_grad else 0,
                   self.total_rows, E, ptr(ws.sorted_rows), ptr(ws.sorted_slots), n, ptr(self.row_begin), ptr(self.row_end),
                   ptr(self.row_stamp), 0, ptr(self.stamp), ptr(ws.dc[0]), H0 + self.K[0] * E, H0, self.K[0], n0,
                   ptr(ws.dc[1]), H0 + self.K[1] * E, H0, self.K[1], n1, self.act, ptr(ws.piece_ws), ws.piece_ws.numel(),
                   ptr(self.hyper), self.adam_mode if apply_adam else TT_ADAM_DENSE_EXACT, 0, 0, 0, s, n_kernels=2)

    def backward(self, ws: _Ws, dloss: float = 1.0):
        self.loss_backward(ws, dloss)
        if self.two_streams:
            main = torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.tower_backward(1, ws)
            self.tower_backward(0, ws)
            main.wait_stream(self._side)
        else:
            self.tower_backward(0, ws)
            self.tower_backward(1, ws)

    # ------------------------------------------------------------------ optimizer
    def hyper_advance(self, lr: float, warmup_steps: int, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        self._call("tt_hyper_advance", ptr(self.hyper), lr, warmup_steps, betas[0], betas[1], eps, weight_decay,
                   self.rng_seed, self._stream())

    def dense_adam(self):
        self._call("tt_dense_adam", ptr(self.dense_w), ptr(self.dense_g), ptr(self.dense_m), ptr(self.dense_v),
                   self.layout.total, ptr(self.hyper), self._stream())

    def fused_step(self, notice_dense, notice_values, company_dense, company_values, lr=1e-3, warmup_steps=0,
                   betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, layout=None, masks=None):
        """hyper_advance -> forward -> backward -> sparse-dedup + Adam on tables -> Adam on dense. Returns out5."""
        self.hyper_advance(lr, warmup_steps, betas, eps, weight_decay)
        out5, ws = self.forward(notice_dense, notice_values, company_dense, company_values, True, layout, masks)
        self.backward(ws)
        self.reduce_dense_grads()
        self.embedding_backward(ws, apply_adam=True, want_dense_grad=False)
        self.dense_adam()
        return out5

    def reduce_dense_grads(self):
        """Single GPU: nothing. Distributed engines all-reduce the flat dense gradient buffer here."""
        return None

    def hyper_host(self) -> dict:
        import struct
        raw = bytes(self.hyper.cpu().numpy().tobytes())
        step, sched, lr, step_size, bc2, b1, b2, eps, wd, _r, seed, rstep = struct.unpack("<qq8fQQ", raw)
        return dict(step=step, sched_step=sched, lr=lr, step_size=step_size, bc2_sqrt=bc2, beta1=b1, beta2=b2, eps=eps,
                    weight_decay=wd, rng_seed=seed, rng_step=rstep)

    def similarity_matrix(self, ws: _Ws) -> torch.Tensor:
        B = ws.B
        S = torch.empty(B, B, dtype=torch.float32, device=self.device)
        n, c = ws.e[0].float(), ws.e[1].float()  # eval / debugging only: the train step never stores the logits
        self._call("tt_similarity_matrix", ptr(n), ptr(c), B, B, self.cfg.final_dim, TT_F32,
                   1.0 / self.cfg.temperature, ptr(S), B, self._stream())
        return S
