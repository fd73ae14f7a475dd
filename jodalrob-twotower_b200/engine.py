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

import os

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native
from .native import TT_ADAM_DENSE_EXACT, TT_ADAM_DENSE_UNTOUCHED, TT_ADAM_SPARSE, TT_F32, TT_LAYOUT_KEY_MAJOR, TT_LAYOUT_SAMPLE_MAJOR, ptr

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LOG2E = 1.4426950408889634


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
    SMALL_TABLE_FLOATS = 2048  # tables with V*E <= this are "small": one-hot dense reduction + dense Adam, never sorted

    def __init__(self, cfg: TwoTowerConfig, device, optimizer_mode: str = "dense_exact", value_layout: str = "sample_major",
                 dtype: str = "fp32", small_tables: bool = True):
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
        if optimizer_mode not in ("dense_exact", "sparse", "lazy_exact"):
            raise ValueError(f"optimizer_mode must be 'dense_exact', 'lazy_exact' or 'sparse', got {optimizer_mode}")
        # lazy_exact: the reference's dense Adam semantics (every row moves every step) without streaming every row every step —
        # rows are caught up, bit for bit, when they are next read or updated (tt_lazy_ctx in include/twotower_b200.h)
        self.lazy = optimizer_mode == "lazy_exact"
        self.adam_mode = {"dense_exact": TT_ADAM_DENSE_EXACT, "sparse": TT_ADAM_SPARSE, "lazy_exact": native.TT_ADAM_LAZY_EXACT}[optimizer_mode]
        # Split exact-dense table update of the fused step (TT_SPLIT_ADAM=1): rows the batch does not touch (g = wd*w) are updated
        # right behind the sort-dedup, underneath the towers; only the touched rows wait for the gradients. Bit-identical to the
        # one-kernel update. Measured on B200 at the README-default size: 376 us per step against 361 us unsplit — the extra kernel
        # lands next to the first sim-CE pass, whose persistent CTAs then find SMs occupied — so it is off by default.
        self.split_dense_adam = os.environ.get("TT_SPLIT_ADAM", "0") != "0"
        # TT_SPLIT_ADAM=2: the untouched-rows half runs behind the logits passes instead, underneath the towers' backward kernels
        # (64-CTA kernels that leave SMs free), so that only the touched rows are left for the end of the step
        self.split_late = os.environ.get("TT_SPLIT_ADAM", "0") == "2"
        # TT_HYPER_HELPER=1: optimizer scalars advanced on a helper stream (was a gain with the per-stage tower kernels; with the fused
        # kernels measured 295.1 us per step against 289.6 us with the plain in-order launch) — off
        self.hyper_on_helper = os.environ.get("TT_HYPER_HELPER", "0") != "0"
        self.defer_metrics = os.environ.get("TT_DEFER_METRICS", "1") != "0"
        # TT_ONE_CAST=1: bf16 weight mirror + both dense inputs cast in one launch at the start of the step. Measured on B200:
        # identical loss bit for bit, but 361.6 us per step against 356.6 us with the three separate casts (the single launch
        # sits in front of both lookups and both projection GEMMs) — off.
        self.one_cast = os.environ.get("TT_ONE_CAST", "0") != "0"
        self._metrics_stream = torch.cuda.Stream(self.device)
        self._arm_split = False
        self._in_fused = False
        self.default_layout = value_layout
        E = cfg.emb_dim
        if E % 4 != 0 or cfg.hidden[0] % 4 != 0:
            raise ValueError("emb_dim and tower_hidden_dims[0] must be multiples of 4 (128-bit vector path)")
        dev = self.device
        self.K = (len(cfg.notice_keys), len(cfg.company_keys))
        self.layout = DenseLayout(cfg)
        self._plan_tables(small_tables and optimizer_mode in ("dense_exact", "lazy_exact"))
        P_tot, RL = self.P_tot, self.total_rows
        # ONE flat fp32 allocation: [dense params | small tables | large-table arena]; Adam moments mirror it
        self.flat_w = torch.zeros(P_tot + RL * E, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros_like(self.flat_w)
        self.flat_v = torch.zeros_like(self.flat_w)
        self.dense_w, self.dense_m, self.dense_v = self.flat_w[:P_tot], self.flat_m[:P_tot], self.flat_v[:P_tot]
        self.dense_g = torch.zeros(P_tot, dtype=torch.float32, device=dev)
        self.arena = self.flat_w[P_tot:].view(RL, E)
        self.arena_m, self.arena_v = self.flat_m[P_tot:].view(RL, E), self.flat_v[P_tot:].view(RL, E)
        self.arena_grad: Optional[torch.Tensor] = None
        self.dense_w16 = torch.zeros(self.layout.total, dtype=torch.bfloat16, device=dev) if dtype == "bf16" else None
        nblk = len(cfg.hidden) - 1
        # Fused tower kernels (tower_fused.cu): bf16, one hidden block, widths the kernels are instantiated for — else the
        # per-stage kernels. TT_FUSED_TOWER=0 forces the per-stage path (A/B measurements, tests of both paths).
        self.fused_tower = bool(
            dtype == "bf16" and nblk == 1 and os.environ.get("TT_FUSED_TOWER", "1") != "0"
            and all(self.lib.query("tt_tower_fused_supported", cfg.hidden[0], cfg.hidden[1], cfg.final_dim,
                                   cfg.hidden[0] + len(cfg.tower(t)[1]) * E) for t in range(2)))
        self.bn_rm = [[torch.zeros(h, device=dev) for h in cfg.hidden[1:]] for _ in range(2)]
        self.bn_rv = [[torch.ones(h, device=dev) for h in cfg.hidden[1:]] for _ in range(2)]
        self.bn_nbt = [[torch.zeros((), dtype=torch.int64, device=dev) for _ in range(nblk)] for _ in range(2)]
        self.hyper = torch.zeros(native.HYPER_BYTES, dtype=torch.uint8, device=dev)
        self.stamp = torch.zeros(1, dtype=torch.int64, device=dev)
        self.row_begin = torch.zeros(max(self.total_rows, 1), dtype=torch.int32, device=dev)
        self.row_end = torch.zeros(max(self.total_rows, 1), dtype=torch.int32, device=dev)
        self.row_stamp = torch.zeros(max(self.total_rows, 1), dtype=torch.int32, device=dev)
        self._lazy_offset = 0      # -1 while a train step is in flight (its own update of the looked-up rows is still to come)
        self._steps_since_flush = 0
        if self.lazy:
            import struct
            self.hist_len = int(os.environ.get("TT_LAZY_HIST", "65536"))
            if self.hist_len & (self.hist_len - 1):
                raise ValueError("TT_LAZY_HIST must be a power of two")
            self.row_last = torch.zeros(max(self.total_rows, 1), dtype=torch.int32, device=dev)
            self.lazy_hist = torch.zeros(2 * self.hist_len, dtype=torch.float32, device=dev)
            raw = struct.pack("<5Qq1Q", self.arena.data_ptr(), self.arena_m.data_ptr(), self.arena_v.data_ptr(), self.row_last.data_ptr(),
                              self.lazy_hist.data_ptr(), self.hist_len, self.hyper.data_ptr())
            self.lazy_ctx = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
        self._ws: Dict[int, _Ws] = {}
        self.rng_seed = 0x5EED
        self.rng_stream_base = 0  # distributed engines offset it by rank so ranks draw different dropout masks
        # the two towers are independent until the loss: run the company tower on a side stream (fork/join with events;
        # inside a CUDA graph capture this becomes two parallel branches)
        self.two_streams = True
        self.sharded_lookup = False  # DistributedEngine: the lookup is one exchange for both towers
        # SyncBN leaves the GLOBAL gamma/beta gradient sums on every rank; the dense-gradient all-reduce must count them once:
        # distributed engines set 1.0 on rank 0 and 0.0 elsewhere
        self.bn_affine_grad_scale = 1.0
        # Streams on the critical path (the towers' data-gradient chains, the table update) get high priority, so that their
        # CTAs are placed before those of the weight-gradient helper streams whenever SMs free up
        import os as _os
        hi = 0 if _os.environ.get("TT_NO_PRIO") else -1
        self._crit_stream = torch.cuda.Stream(self.device, priority=-1) if _os.environ.get("TT_CRIT_PRIO", "0") == "1" else None
        self._side = torch.cuda.Stream(self.device, priority=hi)
        self._hi0 = torch.cuda.Stream(self.device, priority=hi)
        self._sort_stream = torch.cuda.Stream(self.device)  # the sort has the whole forward + loss to hide under: normal priority
        self._adam_stream = torch.cuda.Stream(self.device, priority=hi)
        # two helper streams per tower: the weight gradients of a tower's layers are independent of each other
        self._pre_streams = [torch.cuda.Stream(self.device, priority=hi) for _ in range(2)]
        self._wg_streams = [[torch.cuda.Stream(self.device) for _ in range(4)] for _ in range(2)]  # per tower: 3 weight-gradient chains + small tables
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

    def _alloc_rows(self, t: int, k: int, vocab: int, small: bool) -> int:
        """Rows this process allocates for table (t,k). Distributed engines shard the large tables."""
        return vocab

    def _plan_tables(self, use_small: bool):
        """Split the tables into classes and lay everything out in one flat buffer (see module docstring)."""
        import numpy as np
        cfg, E, dev = self.cfg, self.cfg.emb_dim, self.device
        H0 = cfg.hidden[0]
        unit = E * 64 // math.gcd(E, 64)
        self.P_dense = (self.layout.total + unit - 1) // unit * unit
        self.is_small, self.small_off, self.large_base, self.alloc_rows = [], [], [], []
        s_off, l_row = 0, 0
        large_rank, colmaps, small_tabs, small_floats, small_region = [], [], [], [], []
        self.max_small_vocab = 1
        for t in range(2):
            _, keys, vocab, _ = cfg.tower(t)
            sm, so, lb, ar, rank, cm, tabs = [], [], [], [], [], [], []
            region0 = s_off
            nl = 0
            for k, v in enumerate(vocab):
                small = bool(use_small and v * E <= self.SMALL_TABLE_FLOATS)
                rows = self._alloc_rows(t, k, int(v), small)
                sm.append(small)
                ar.append(rows)
                if small:
                    so.append(s_off)
                    lb.append(-1)
                    rank.append(-1)
                    tabs.append((k, int(v), H0 + k * E, s_off - region0))
                    s_off += rows * E
                    self.max_small_vocab = max(self.max_small_vocab, int(v))
                else:
                    so.append(-1)
                    lb.append(l_row)
                    rank.append(nl)
                    cm.append(H0 + k * E)
                    nl += 1
                    l_row += rows
            self.is_small.append(sm); self.small_off.append(so); self.large_base.append(lb); self.alloc_rows.append(ar)
            large_rank += rank
            colmaps.append(cm)
            small_region.append(region0)
            small_floats.append(s_off - region0)
            arr = np.zeros(len(tabs), dtype=np.dtype([("key", "<i4"), ("vocab", "<i4"), ("col", "<i4"), ("pad", "<i4"), ("g_off", "<i8")]))
            for i, (k, v, col, go) in enumerate(tabs):
                arr[i] = (k, v, col, 0, go)
            small_tabs.append(torch.from_numpy(arr.view(np.uint8).copy()).to(dev) if len(tabs) else None)
        self.S_floats = s_off
        self.P_tot = (self.P_dense + s_off + unit - 1) // unit * unit
        self.total_rows = l_row  # rows of the large-table arena (sort keys live in [0, total_rows))
        self.KL = (len(colmaps[0]), len(colmaps[1]))
        self.n_small = (0 if small_tabs[0] is None else small_tabs[0].numel() // 24, 0 if small_tabs[1] is None else small_tabs[1].numel() // 24)
        self.small_tabs, self.small_floats, self.small_region = small_tabs, small_floats, small_region
        self.large_rank = torch.tensor(large_rank, dtype=torch.int32, device=dev)
        self.colmap = [torch.tensor(c if c else [0], dtype=torch.int32, device=dev) for c in colmaps]
        base, vocab_all = [], []
        for t in range(2):
            _, keys, vocab, _ = cfg.tower(t)
            for k, v in enumerate(vocab):
                if self.is_small[t][k]:
                    base.append((self.P_dense + self.small_off[t][k]) // E)
                else:
                    base.append(self.P_tot // E + self.large_base[t][k])
                vocab_all.append(int(v))
        self.row_base_host = base
        self.row_base = torch.tensor(base, dtype=torch.int64, device=dev)  # rows relative to flat_w
        self.vocab = torch.tensor(vocab_all, dtype=torch.int32, device=dev)

    def dense(self, name: str, which: str = "w") -> torch.Tensor:
        return self.layout.view({"w": self.dense_w, "g": self.dense_g, "m": self.dense_m, "v": self.dense_v}[which], name)

    def table_view(self, t: int, k: int, which: str = "w") -> torch.Tensor:
        """[rows, E] view of table (t,k): weights / Adam moments / gradient (gradient of a large table needs arena_grad)."""
        E = self.cfg.emb_dim
        rows = self.alloc_rows[t][k]
        if self.is_small[t][k]:
            buf = {"w": self.dense_w, "m": self.dense_m, "v": self.dense_v, "g": self.dense_g}[which]
            o = self.P_dense + self.small_off[t][k]
            return buf[o:o + rows * E].view(rows, E)
        buf = {"w": self.arena, "m": self.arena_m, "v": self.arena_v, "g": self.arena_grad}[which]
        lo = self.large_base[t][k]
        return buf[lo:lo + rows]

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
        n_tot = B * (self.KL[0] + self.KL[1])
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
                    ws.x16, ws.dz16, ws.x16_owned = [], [], []
                ws.x16.append(torch.empty(B, dense_dim, **act))
                ws.x16_owned.append(True)
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
        if self.fused_tower:
            nt = int(self.lib.query("tt_linear_stats_tiles", B))
            ws.n_tiles = nt
            H0_, H1_, F_ = cfg.hidden[0], cfg.hidden[1], cfg.final_dim
            ws.stat_part = [torch.zeros(nt * 2 * H1_, **f32) for _ in range(2)]
            ws.bn_part = [torch.zeros(nt * 2 * H1_, **f32) for _ in range(2)]
            ws.dz_colpart = [torch.zeros(nt * F_, **f32) for _ in range(2)]
            ws.dh_colpart = [torch.zeros(nt * H1_, **f32) for _ in range(2)]
            ws.dc_colpart = [torch.zeros(nt * H0_, **f32) for _ in range(2)]
        # one scratch per tower so the two towers may run on different streams
        ws.bn_ws = [torch.empty(max(max_bn, 256), dtype=torch.uint8, device=dev) for _ in range(2)]
        ws.bn_sums = [torch.zeros(2 * max(cfg.hidden[1:] + [1]), dtype=torch.float64, device=dev) for _ in range(2)]
        ws.wg_ws = [[torch.empty(max_wg, dtype=torch.uint8, device=dev) for _ in range(4)] for _ in range(2)]
        ws.sim_ws = torch.empty(q("tt_simce_workspace_bytes", B, B, cfg.final_dim), dtype=torch.uint8, device=dev)
        ws.row_sumexp = torch.empty(B, **f32)
        ws.col_sumexp = torch.empty(B, **f32)
        ws.row_sum_s = torch.empty(B, **f32)
        ws.col_sum_s = torch.empty(B, **f32)
        ws.diag = torch.empty(B, **f32)
        ws.row_argmax = torch.empty(B, dtype=torch.int32, device=dev)
        ws.inv_row = torch.empty(B, **f32)
        ws.inv_col = torch.empty(B, **f32)
        # deferred-metrics forward (fused step, bf16): its own workspace (the backward calls reuse sim_ws while the metrics stream
        # still reads the forward partials) and scratch outputs for the reciprocals tt_simce_finalize recomputes there
        split_bytes = q("tt_simce_split_workspace_bytes", B, B, cfg.final_dim) if self.emb_tt != TT_F32 else 0
        ws.sim_ws_fwd = torch.empty(max(split_bytes, 16), dtype=torch.uint8, device=dev)
        ws.inv_scratch = [torch.empty(B, **f32), torch.empty(B, **f32)]
        self._alloc_sim3(ws, B, B)
        ws.partial = torch.zeros(128, dtype=torch.float64, device=dev)  # 16 block partials x 8
        ws.out5 = torch.zeros(5, **f32)
        ws.sort_ws = torch.empty(q("tt_ebc_sort_workspace_bytes", n_tot, max(self.total_rows, 1)), dtype=torch.uint8, device=dev)
        ws.piece_ws = torch.empty(q("tt_ebc_piece_workspace_bytes", n_tot, E), dtype=torch.uint8, device=dev)
        ws.small_ws = [torch.empty(q("tt_ebc_small_workspace_bytes", B, max(self.small_floats[t], 1)), dtype=torch.uint8, device=dev)
                       for t in range(2)]
        self._ws[B] = ws
        return ws

    def _alloc_sim3(self, ws: _Ws, Bx: int, By: int):
        """Buffers of the three-pass logits step (bf16 / tcgen05, fused step): es = tower outputs pre-scaled by log2(e)/T (the X
        operand of every pass), xs = the other side's rows scaled by g / sum_exp (the Ys operand of the gradient passes)."""
        cfg, dev = self.cfg, self.device
        ws.sim3_ws = None
        if self.emb_tt == TT_F32 or os.environ.get("TT_SIMCE_PASSES", "3") == "4":
            return
        bf = dict(dtype=torch.bfloat16, device=dev)
        ws.es = [torch.empty(Bx, cfg.final_dim, **bf) for _ in range(2)]
        ws.xs = [torch.empty(By, cfg.final_dim, **bf) for _ in range(2)]
        ws.sim3_ws = torch.empty(self.lib.query("tt_simce3_workspace_bytes", Bx, By, cfg.final_dim), dtype=torch.uint8, device=dev)
        ws.pending3 = False

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
            self._call("tt_linear_wgrad_bf16", ptr(dy), ldy, ptr(x), ldx, ptr(self.dense(wname, "g")),
                       ptr(self.dense(bname, "g")) if bname else 0, M, N, K, ptr(wsbuf), wsbuf.numel(), s,
                       n_kernels=4 if bname else 2)

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
            rows_off = 0 if t == 0 else B * self.KL[0]
            self.profile_tag = self.cfg.tower(t)[0]
            if self.lazy and self.total_rows > 0:
                self._call("tt_ebc_forward_mixed_lazy", ptr(self.flat_w), self.row_base.data_ptr() + 8 * k0, self.vocab.data_ptr() + 4 * k0,
                           ptr(v), B, K, E, layout, ptr(ws.c[t]), self.act, H0 + K * E, H0, self.large_rank.data_ptr() + 4 * k0,
                           self.KL[t], self.P_tot // E, ws.rows.data_ptr() + 4 * rows_off, ptr(self.lazy_ctx), self._lazy_offset, s)
                continue
            self._call("tt_ebc_forward_mixed", ptr(self.flat_w), self.row_base.data_ptr() + 8 * k0, self.vocab.data_ptr() + 4 * k0,
                       ptr(v), B, K, E, layout, ptr(ws.c[t]), self.act, H0 + K * E, H0, self.large_rank.data_ptr() + 4 * k0,
                       self.KL[t], self.P_tot // E, ws.rows.data_ptr() + 4 * rows_off, s)
        ws.values, ws.layout = list(values), layout

    def tower_prefix(self, t: int, ws: _Ws, dense: torch.Tensor, cast_done: bool = False):
        """The part of a tower that does not need the embedding lookup: the dense projection into columns [0, H0) of the
        concat buffer. The train step runs it on a helper stream underneath the lookup."""
        cfg, B, s = self.cfg, ws.B, self._stream()
        name, keys, _, dense_dim = cfg.tower(t)
        self.profile_tag = name
        H0 = cfg.hidden[0]
        ld = H0 + len(keys) * cfg.emb_dim
        pre = f"{name}_tower."
        dense = dense.contiguous()
        if self.act != TT_F32 and dense.dtype == torch.bfloat16 and dense.data_ptr() % 16 == 0:
            # already in the compute dtype (GraphedFusedStep(compact_inputs=True) stages the dense features as bf16: half the bytes
            # over PCIe, no cast kernel; same bits as the device-side round-to-nearest-even cast): the weight gradient reads it too
            ws.x16[t] = dense
            ws.x16_owned[t] = False
        else:
            if dense.dtype != torch.float32:
                dense = dense.float()
            if self.act != TT_F32:
                if ws.x16[t].shape != dense.shape or ws.x16[t].data_ptr() == dense.data_ptr() or not ws.x16_owned[t]:
                    ws.x16[t] = torch.empty(B, dense_dim, dtype=self.act_torch, device=self.device)
                    ws.x16_owned[t] = True
                if not cast_done:
                    self._call("tt_cast_bf16", ptr(dense), ptr(ws.x16[t]), B * dense_dim, s)
                dense = ws.x16[t]
        self._lin_fwd(dense, dense_dim, pre + "dense_projection.weight", pre + "dense_projection.bias", ws.c[t], ld, False, B, H0,
                      dense_dim, 0)

    def tower_forward(self, t: int, ws: _Ws, dense: torch.Tensor, values: Optional[torch.Tensor], train: bool,
                      layout: int = TT_LAYOUT_SAMPLE_MAJOR, masks: Optional[Sequence[torch.Tensor]] = None,
                      skip_prefix: bool = False):
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
        if not skip_prefix:
            # stand-alone tower call (TwoTowerModel.forward, get_*_embeddings, predict_batch): nobody refreshed the bf16 weight
            # mirror for us — after load_state_dict / an optimizer step it is stale (all zeros on a fresh engine)
            self.refresh_bf16_weights()
            self.tower_prefix(t, ws, dense)
            self.profile_tag = name
        if self.fused_tower and train:
            # two kernels: [Linear1 + ReLU + per-tile column statistics] -> [BN finalize/apply + dropout -> Linear2 + bias -> L2 norm]
            h, F = cfg.hidden[1], cfg.final_dim
            self._call("tt_linear_forward_bf16_stats", ptr(ws.c[t]), ld, ptr(self._w16(pre + "mlp.0.weight")), ptr(self.dense(pre + "mlp.0.bias")),
                       ptr(ws.h[t][0]), h, native.TT_BF16, B, h, ld, 1, ptr(ws.stat_part[t]), s)
            count, sums = self.bn_sync(ws.stat_part[t], ws.n_tiles, 2 * h, B, t, ws.bn_sums[t])
            p = cfg.dropout
            mask = masks[0] if (masks is not None and p > 0) else None
            es = ws.es[t] if getattr(ws, "sim3_ws", None) is not None else None
            self._call("tt_tower_tail_forward", ptr(ws.h[t][0]), B, h, F, ptr(ws.stat_part[t]), ws.n_tiles, ptr(sums), count,
                       ptr(self.dense(pre + "mlp.2.weight")), ptr(self.dense(pre + "mlp.2.bias")), BN_EPS, BN_MOMENTUM,
                       ptr(ws.mean[t][0]), ptr(ws.rstd[t][0]), ptr(self.bn_rm[t][0]), ptr(self.bn_rv[t][0]), ptr(self.bn_nbt[t][0]), p,
                       ptr(mask), ptr(self.hyper), self.rng_stream_base + t * 16, ptr(ws.y[t][0]), ptr(self._w16(pre + "mlp.4.weight")),
                       ptr(self.dense(pre + "mlp.4.bias")), ptr(ws.e[t]), ptr(es), LOG2E / cfg.temperature, ptr(ws.nrm[t]), s)
            return ws.e[t]
        x, width, idx = ws.c[t], ld, 0
        for bi, h in enumerate(cfg.hidden[1:]):
            self._lin_fwd(x, width, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", ws.h[t][bi], h, False, B, h, width, 1)
            gamma, beta = self.dense(pre + f"mlp.{idx + 2}.weight"), self.dense(pre + f"mlp.{idx + 2}.bias")
            if train:
                self._call("tt_bn_stats", ptr(ws.h[t][bi]), h, B, h, self.act, ptr(ws.bn_sums[t]), ptr(ws.bn_ws[t]),
                           ws.bn_ws[t].numel(), s, n_kernels=2)
                count = self.bn_allreduce(ws.bn_sums[t], 2 * h, B, t)
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
        if getattr(ws, "sim3_ws", None) is not None:
            # unit rows AND rows pre-scaled by log2(e)/T (rounded once from fp32): the logits passes then need no multiply per logit
            self._call("tt_l2norm_forward_scaled", ptr(ws.z[t]), B, F, ptr(ws.e[t]), ptr(ws.es[t]), LOG2E / cfg.temperature,
                       ptr(ws.nrm[t]), s)
        else:
            self._call("tt_l2norm_forward", ptr(ws.z[t]), B, F, ptr(ws.e[t]), self.emb_tt, ptr(ws.nrm[t]), s)
        return ws.e[t]

    def bn_allreduce(self, sums: torch.Tensor, n: int, local_count: int, tower: int = 0) -> int:
        """Single GPU: nothing to do. Distributed engines all-reduce `sums[:n]` and return the global count."""
        return local_count

    def bn_sync(self, part: torch.Tensor, n_tiles: int, n: int, local_count: int, tower: int, sums: torch.Tensor):
        """Fused tower kernels: per-tile column partials `part` [n_tiles][n] of this rank's rows. Single GPU: the consumer folds
        them itself -> (count, None). Distributed engines fold them into `sums` (f64 [n]), all-reduce, and return (global count,
        sums)."""
        return local_count, None

    def loss_forward(self, ws: _Ws):
        cfg, B, s = self.cfg, ws.B, self._stream()
        self.profile_tag = ""
        inv_t = 1.0 / cfg.temperature
        css = ptr(ws.col_sum_s) if cfg.label_smoothing > 0 else 0  # column sums of S are only needed for label smoothing
        ws.pending3 = False
        if self._in_fused and getattr(ws, "sim3_ws", None) is not None:
            # Fused step, tensor-core path: THREE logits passes per step instead of four. Here: P1 (row sums of exp) and its
            # reduce, which also writes the notice rows scaled by g / row_sum — the second MMA operand of P2. The column sums come
            # out of P2 (loss_backward), the loss value and the metrics after that, off the critical path.
            F, nb = cfg.final_dim, ws.sim3_ws.numel()
            ws.gscale3 = 0.5 / B * inv_t
            self._call("tt_simce3_rows", ptr(ws.es[0]), ptr(ws.e[1]), B, B, F, inv_t, 0, ptr(ws.diag), ptr(ws.sim3_ws), nb, s)
            self._call("tt_simce3_rows_finish", ptr(ws.sim3_ws), nb, B, B, F, inv_t, ptr(ws.row_sumexp), ptr(ws.inv_row), ptr(ws.e[0]),
                       ws.gscale3, ptr(ws.xs[0]), s)
            ws.pending3 = True
            return ws.out5
        if (self.emb_tt != TT_F32 and self._in_fused and self.profile is None and self.defer_metrics and self.two_streams
                and ws.sim_ws_fwd.numel() > 16):
            # Fused step, tensor-core path: only the row/column sums of exp (and their reciprocals) are on the critical path — the
            # backward needs nothing else. sum_j S_ij, the arg-max resolution, the loss partials and the five scalars run on a side
            # stream from the partials left in the workspace and are joined at the end of the step (_end_of_step).
            main = torch.cuda.current_stream(self.device)
            self._call("tt_simce_forward_sums", ptr(ws.e[0]), ptr(ws.e[1]), B, B, cfg.final_dim, self.emb_tt, inv_t, 0,
                       ptr(ws.row_sumexp), ptr(ws.col_sumexp), ptr(ws.inv_row), ptr(ws.inv_col), ptr(ws.diag), ptr(ws.sim_ws_fwd),
                       ws.sim_ws_fwd.numel(), s, n_kernels=4)

            def metrics():  # launched by fused_step BEHIND the backward logits passes (see there)
                ms_ = self._metrics_stream
                ms_.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(ms_):
                    ms = ms_.cuda_stream
                    self._call("tt_simce_forward_metrics", ptr(ws.e[0]), ptr(ws.e[1]), B, B, cfg.final_dim, self.emb_tt, inv_t,
                               ptr(ws.row_sum_s), css, ptr(ws.row_argmax), ptr(ws.sim_ws_fwd), ws.sim_ws_fwd.numel(), ms, n_kernels=2)
                    self._call("tt_simce_finalize", ptr(ws.row_sumexp), ptr(ws.col_sumexp), ptr(ws.row_sum_s), css, ptr(ws.diag),
                               ptr(ws.row_argmax), B, B, 0, B, inv_t, cfg.label_smoothing, ptr(ws.inv_scratch[0]),
                               ptr(ws.inv_scratch[1]), ptr(ws.partial), ms)
                    self._call("tt_simce_scalars", ptr(ws.partial), B, ptr(ws.out5), ms)
                    ws.metrics_done = torch.cuda.Event()
                    ws.metrics_done.record(ms_)

            ws.deferred_metrics = metrics
            return ws.out5
        self._call("tt_simce_forward", ptr(ws.e[0]), ptr(ws.e[1]), B, B, cfg.final_dim, self.emb_tt, inv_t, 0, ptr(ws.row_sumexp),
                   ptr(ws.col_sumexp), ptr(ws.row_sum_s), css, ptr(ws.diag), ptr(ws.row_argmax),
                   ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=4)
        self._call("tt_simce_finalize", ptr(ws.row_sumexp), ptr(ws.col_sumexp), ptr(ws.row_sum_s), css,
                   ptr(ws.diag), ptr(ws.row_argmax), B, B, 0, B, inv_t, cfg.label_smoothing, ptr(ws.inv_row),
                   ptr(ws.inv_col), ptr(ws.partial), s)
        self._call("tt_simce_scalars", ptr(ws.partial), B, ptr(ws.out5), s)
        return ws.out5

    def forward(self, notice_dense, notice_values, company_dense, company_values, train: bool = True,
                layout: Optional[str] = None, masks=None, sort_on_side: bool = False) -> Tuple[torch.Tensor, _Ws]:
        B = notice_dense.shape[0]
        if company_dense.shape[0] != B:
            raise ValueError(f"notice/company batch sizes differ: {B} vs {company_dense.shape[0]}")
        self._check_inputs(0, notice_dense, notice_values, B)
        self._check_inputs(1, company_dense, company_values, B)
        for t, v in ((0, notice_values), (1, company_values)):
            if v.numel() != B * self.K[t] * getattr(self, "bag_len", 1):
                raise ValueError(f"kjt values hold {v.numel()} ids, expected B*K = {B * self.K[t] * getattr(self, 'bag_len', 1)}")
        lay = TT_LAYOUT_KEY_MAJOR if (layout or self.default_layout) == "key_major" else TT_LAYOUT_SAMPLE_MAJOR
        ws = self.workspace(B)
        ws.inputs = (notice_dense.contiguous(), company_dense.contiguous())
        main = torch.cuda.current_stream(self.device)
        overlap = self.two_streams and self.profile is None
        one_cast = (overlap and self.one_cast and self.act != TT_F32 and all(d.dtype == torch.float32 for d in ws.inputs))
        if one_cast:  # the bf16 weight mirror and both dense inputs in ONE launch: every extra root node of the graph costs
            nd_, cd_ = ws.inputs  # 5-9 us of launch ramp at the start of a replay
            self._call("tt_cast_bf16_multi", ptr(self.dense_w), ptr(self.dense_w16), self.layout.total, ptr(nd_), ptr(ws.x16[0]),
                       nd_.numel(), ptr(cd_), ptr(ws.x16[1]), cd_.numel(), 0, 0, 0, self._stream())
        else:
            self.refresh_bf16_weights()
        if overlap:  # the dense projections do not need the lookup: run them underneath it on helper streams
            pre = self._pre_streams  # high priority: their GEMM CTAs take the SM slots the lookup's first wave frees
            for t, d in ((0, notice_dense), (1, company_dense)):
                pre[t].wait_stream(main)
                with torch.cuda.stream(pre[t]):
                    self.tower_prefix(t, ws, d, cast_done=one_cast)
        split = overlap and not self.sharded_lookup
        if split:  # the two towers' lookups side by side (the company tower continues on `_side` anyway)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.embedding_forward(ws, [None, company_values], lay)
            self.embedding_forward(ws, [notice_values, None], lay)
            ws.values = [notice_values, company_values]
        else:
            self.embedding_forward(ws, [notice_values, company_values], lay)
        hd = getattr(self, "_hyper_done", None)
        if hd is not None:  # the optimizer / dropout scalars of this step (advanced on a helper stream by fused_step)
            main.wait_event(hd)
            self._side.wait_event(hd)
        if sort_on_side:  # the index side of the sparse backward only needs the looked-up rows
            self._sort_stream.wait_stream(main)
            if split:
                self._sort_stream.wait_stream(self._side)
            with torch.cuda.stream(self._sort_stream):
                self.embedding_sort(ws)
        m0, m1 = (None, None) if masks is None else (masks[0], masks[1])
        if self.two_streams:
            if not split:
                self._side.wait_stream(main)
            if overlap:
                main.wait_stream(pre[0])
                self._side.wait_stream(pre[1])
            with torch.cuda.stream(self._side):
                self.tower_forward(1, ws, company_dense, None, train, lay, masks=m1, skip_prefix=overlap)
            self.tower_forward(0, ws, notice_dense, None, train, lay, masks=m0, skip_prefix=overlap)
            main.wait_stream(self._side)
        else:
            self.tower_forward(0, ws, notice_dense, None, train, lay, masks=m0)
            self.tower_forward(1, ws, company_dense, None, train, lay, masks=m1)
        out5 = self.loss_forward(ws)
        ws.train = train
        ws.masks = masks
        ws.dc_ready = [None, None]
        return out5, ws

    # ------------------------------------------------------------------ backward
    def loss_backward(self, ws: _Ws, dloss: float = 1.0):
        cfg, B, s = self.cfg, ws.B, self._stream()
        self.profile_tag = ""
        inv_t = 1.0 / cfg.temperature
        gscale = dloss * 0.5 / B * inv_t
        F = cfg.final_dim
        if getattr(ws, "pending3", False):
            ws.pending3 = False
            nb, ls, g = ws.sim3_ws.numel(), cfg.label_smoothing, ws.gscale3
            css = ptr(ws.col_sum_s) if ls > 0 else 0
            # P2: the transposed problem (rows = company): column sums of exp on the fly + the raw halves of dC; its reduce applies the
            # row scale, writes col_sumexp / inv_col and the company rows scaled by g / col_sum — the second operand of P3
            self._call("tt_simce3_grad", ptr(ws.es[1]), ptr(ws.e[0]), ptr(ws.xs[0]), B, B, F, inv_t, 0, 0, g, ls, ptr(ws.sim3_ws), nb, s)
            self._call("tt_simce3_grad_finish", ptr(ws.sim3_ws), nb, B, B, F, inv_t, 0, ptr(ws.e[0]), 1, g, ls, B, ptr(ws.de[1]),
                       ptr(ws.col_sumexp), ptr(ws.inv_col), ptr(ws.e[1]), ptr(ws.xs[1]), s)
            # P3: dN, row scale known from P1
            self._call("tt_simce3_grad", ptr(ws.es[0]), ptr(ws.e[1]), ptr(ws.xs[1]), B, B, F, inv_t, 0, ptr(ws.inv_row), g, ls,
                       ptr(ws.sim3_ws), nb, s)
            self._call("tt_simce3_grad_finish", ptr(ws.sim3_ws), nb, B, B, F, inv_t, 0, ptr(ws.e[1]), 0, g, ls, B, ptr(ws.de[0]), 0, 0,
                       0, 0, s)

            def metrics(ms):  # loss value + metrics from the P1 partials and the sums
                self._call("tt_simce3_metrics", ptr(ws.es[0]), ptr(ws.e[1]), B, B, F, ptr(ws.row_sum_s), ptr(ws.row_argmax),
                           ptr(ws.sim3_ws), nb, ms, n_kernels=2)
                if css:
                    self._call("tt_simce3_metrics", ptr(ws.es[1]), ptr(ws.e[0]), B, B, F, css, 0, ptr(ws.sim3_ws), nb, ms, n_kernels=2)
                self._call("tt_simce_finalize", ptr(ws.row_sumexp), ptr(ws.col_sumexp), ptr(ws.row_sum_s), css, ptr(ws.diag),
                           ptr(ws.row_argmax), B, B, 0, B, inv_t, ls, ptr(ws.inv_scratch[0]), ptr(ws.inv_scratch[1]),
                           ptr(ws.partial), ms)
                self._call("tt_simce_scalars", ptr(ws.partial), B, ptr(ws.out5), ms)

            if self.defer_metrics and self.two_streams and self.profile is None:
                def deferred():  # launched by fused_step behind the gradient passes, joined at the end of the step
                    ms_ = self._metrics_stream
                    ms_.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(ms_):
                        metrics(ms_.cuda_stream)
                        ws.metrics_done = torch.cuda.Event()
                        ws.metrics_done.record(ms_)
                ws.deferred_metrics = deferred
            else:
                metrics(s)
            return
        self._call("tt_simce_backward", ptr(ws.e[0]), ptr(ws.e[1]), B, B, F, self.emb_tt, inv_t, 0, ptr(ws.inv_row), ptr(ws.inv_col),
                   gscale, cfg.label_smoothing, B, ptr(ws.de[0]), ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=2)
        self._call("tt_simce_backward", ptr(ws.e[1]), ptr(ws.e[0]), B, B, F, self.emb_tt, inv_t, 0, ptr(ws.inv_col), ptr(ws.inv_row),
                   gscale, cfg.label_smoothing, B, ptr(ws.de[1]), ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=2)

    def tower_backward(self, t: int, ws: _Ws, after_dc=None):
        """Backward of one tower. `after_dc`, if given, is called on this tower's stream as soon as the gradient of the concat
        buffer exists (the fused step hangs the large-table update there: same stream, no cross-stream hand-off on the
        critical path); the rest of this tower (small tables, weight gradients) then goes to the helper streams."""
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        name, keys, _, dense_dim = cfg.tower(t)
        self.profile_tag = name + ".bwd"
        H0, F = cfg.hidden[0], cfg.final_dim
        ld = H0 + len(keys) * E
        pre = f"{name}_tower."
        train = ws.train
        nblk = len(cfg.hidden) - 1
        widths = [ld] + list(cfg.hidden[1:])
        fused = self.fused_tower and train
        bias_of = (lambda n_: None) if fused else (lambda n_: n_)  # fused kernels leave bias gradients as per-tile partials
        if fused:
            # two kernels: [L2-norm backward -> dgrad Linear2 -> dropout mask -> BN-backward sums] -> [BN backward + ReLU mask ->
            # dgrad Linear1 over all columns of the concat buffer + bias-gradient partials]
            h = cfg.hidden[1]
            p = cfg.dropout
            mask = ws.masks[t][0] if (ws.masks is not None and p > 0) else None
            self._call("tt_tower_head_backward", ptr(ws.de[t]), ptr(ws.e[t]), ptr(ws.nrm[t]), B, F, h, ptr(self._w16(pre + "mlp.4.weight")),
                       ptr(ws.dz16[t]), ptr(ws.dz_colpart[t]), ptr(ws.h[t][0]), ptr(ws.mean[t][0]), ptr(ws.rstd[t][0]), p, ptr(mask),
                       ptr(self.hyper), self.rng_stream_base + t * 16, ptr(ws.dy[t][0]), ptr(ws.bn_part[t]), s)
            count, sums = self.bn_sync(ws.bn_part[t], ws.n_tiles, 2 * h, B, t, ws.bn_sums[t])
            self._call("tt_tower_mid_backward", ptr(ws.dy[t][0]), ptr(ws.h[t][0]), B, h, ld, ptr(self._w16(pre + "mlp.0.weight")),
                       ptr(ws.bn_part[t]), ws.n_tiles, ptr(sums), count, ptr(ws.mean[t][0]), ptr(ws.rstd[t][0]),
                       ptr(self.dense(pre + "mlp.2.weight")), ptr(self.dense(pre + "mlp.2.weight", "g")),
                       ptr(self.dense(pre + "mlp.2.bias", "g")), self.bn_affine_grad_scale, ptr(ws.dz_colpart[t]), F,
                       ptr(self.dense(pre + "mlp.4.bias", "g")), ptr(ws.dh[t][0]), ptr(ws.dh_colpart[t]), ptr(ws.dc[t]), H0,
                       ptr(ws.dc_colpart[t]), s)
            wgrads = [(ws.dz16[t], F, ws.y[t][0], h, pre + "mlp.4.weight", None, F, h),
                      (ws.dh[t][0], h, ws.c[t], ld, pre + "mlp.0.weight", None, h, ld)]
        else:
            self._call("tt_l2norm_backward", ptr(ws.de[t]), ptr(ws.e[t]), self.emb_tt, ptr(ws.nrm[t]), B, F, ptr(ws.dz[t]), s)
            idx = 4 * nblk
            xlast = ws.y[t][nblk - 1] if nblk > 0 else ws.c[t]
            wl = widths[-1]
            dz = ws.dz[t]
            if self.act != TT_F32:
                self._call("tt_cast_bf16", ptr(ws.dz[t]), ptr(ws.dz16[t]), B * F, s)
                dz = ws.dz16[t]
            # The chain of data gradients runs first, top to bottom: it is the critical path (at its end the large-table update can
            # start). Every layer's dY stays in its own buffer, so the weight gradients are formed afterwards, off that path.
            wgrads = [(dz, F, xlast, wl, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", F, wl)]
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
                count = self.bn_allreduce(ws.bn_sums[t], 2 * h, B, t)
                self._call("tt_bn_backward_apply", ptr(ws.dy[t][bi]), h, ptr(ws.h[t][bi]), h, B, h, self.act, ptr(mean),
                           ptr(ws.rstd[t][bi]), ptr(gamma), p, ptr(mask), ptr(self.hyper), self.rng_stream_base + t * 16 + bi, ptr(ws.bn_sums[t]),
                           count, 1 if train else 0, ptr(self.dense(pre + f"mlp.{idx + 2}.weight", "g")),
                           ptr(self.dense(pre + f"mlp.{idx + 2}.bias", "g")), self.bn_affine_grad_scale, ptr(ws.dh[t][bi]), h, s)
                xin = ws.y[t][bi - 1] if bi > 0 else ws.c[t]
                win = widths[bi]
                dxin = ws.dy[t][bi - 1] if bi > 0 else ws.dc[t]
                self._lin_dgrad(ws.dh[t][bi], h, pre + f"mlp.{idx}.weight", dxin, win, B, h, win)
                wgrads.append((ws.dh[t][bi], h, xin, win, pre + f"mlp.{idx}.weight", pre + f"mlp.{idx}.bias", h, win))
        self._mark_dc_ready(t, ws)
        cur = torch.cuda.current_stream(self.device)
        fan = self.two_streams and self.profile is None
        inline = after_dc is not None and fan
        joins, used = [], set()
        if after_dc is not None:
            after_dc()
        if inline:  # this stream is busy with the table update: small tables on a helper of their own
            st = self._wg_streams[t][3]
            st.wait_event(ws.dc_ready[t])
            used.add(3)
            with torch.cuda.stream(st):
                self.small_tables_backward(ws, (t,))
        else:
            self.small_tables_backward(ws, (t,))  # needs dc only; the dense Adam at the very end of the step waits for it
        self.profile_tag = name + ".bwd"
        if self.act != TT_F32:
            dense_in = ws.x16[t]
        else:
            dense_in = ws.inputs[t]
            if dense_in.dtype != torch.float32:
                dense_in = dense_in.float()
        wgrads.append((ws.dc[t], ld, dense_in, dense_dim, pre + "dense_projection.weight", bias_of(pre + "dense_projection.bias"), H0,
                       dense_dim))
        # every dY exists now; the layers' weight gradients are independent chains of four small kernels: each gets a helper
        # stream of its own (own scratch each; the README-default tower has three), join at the end
        for i, (dy_, ldy_, x_, ldx_, wn_, bn_, n_, k_) in enumerate(wgrads):
            lane = 1 + i % 3 if fan else 0
            if lane == 0:
                self._lin_wgrad(dy_, ldy_, x_, ldx_, wn_, bn_, B, n_, k_, ws.wg_ws[t][0])
            else:
                st = self._wg_streams[t][lane - 1]
                st.wait_event(ws.dc_ready[t])
                used.add(lane - 1)
                with torch.cuda.stream(st):
                    self._lin_wgrad(dy_, ldy_, x_, ldx_, wn_, bn_, B, n_, k_, ws.wg_ws[t][lane])
        if fused:  # the bias gradients of Linear1 and of the dense projection: one fold of the per-tile partials
            def bias_fold():
                self._call("tt_colpart_reduce2", ptr(ws.dh_colpart[t]), cfg.hidden[1], ptr(self.dense(pre + "mlp.0.bias", "g")),
                           ptr(ws.dc_colpart[t]), H0, ptr(self.dense(pre + "dense_projection.bias", "g")), ws.n_tiles, self._stream())
            if fan:  # behind the small-table reduction on its helper stream (or alone there)
                st = self._wg_streams[t][3]
                if 3 not in used:
                    st.wait_event(ws.dc_ready[t])
                    used.add(3)
                with torch.cuda.stream(st):
                    bias_fold()
            else:
                bias_fold()
        for j in sorted(used):
            ev = torch.cuda.Event()
            ev.record(self._wg_streams[t][j])
            joins.append(ev)
        if inline:
            # this stream still carries the table update: hand the helper joins to the caller, so that the dense-gradient
            # all-reduce / dense Adam can start when the weight gradients are done, next to the table update
            ws.wg_joins = joins
        for ev in joins:
            cur.wait_event(ev)

    def _mark_dc_ready(self, t: int, ws: _Ws):
        """The gradient w.r.t. the tower's concat buffer is complete: everything the large-table update needs from this tower."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        ws.dc_ready[t] = ev

    def small_tables_backward(self, ws: _Ws, towers=(0, 1)):
        """Gradient of the small tables into the flat dense gradient buffer (dense one-hot reduction, deterministic)."""
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        H0 = cfg.hidden[0]
        for t in towers:
            if self.n_small[t] == 0:
                continue
            self.profile_tag = cfg.tower(t)[0]
            gout = self.dense_g.data_ptr() + 4 * (self.P_dense + self.small_region[t])
            self._call("tt_ebc_small_backward", ptr(ws.values[t]), B, self.K[t], ws.layout, ptr(ws.dc[t]), H0 + self.K[t] * E, self.act,
                       E, ptr(self.small_tabs[t]), self.n_small[t], self.max_small_vocab, self.small_floats[t], gout,
                       ptr(ws.small_ws[t]), ws.small_ws[t].numel(), s, n_kernels=2)

    def embedding_sort(self, ws: _Ws):
        """Large tables, index side: sort-by-row dedup of both towers' (row, slot) pairs. Depends only on the forward
        lookup, so the fused step runs it on a side stream underneath the towers and the loss."""
        B, s = ws.B, self._stream()
        n = B * (self.KL[0] + self.KL[1])
        if n == 0 or self.total_rows == 0:
            return
        self.profile_tag = ""
        # the dedup epoch (self.stamp, device counter) is advanced inside the sort's first kernel
        self._call("tt_ebc_sort_rows_advance", ptr(ws.rows), 0, n, self.total_rows, ptr(ws.sort_ws), ws.sort_ws.numel(),
                   ptr(ws.sorted_rows), ptr(ws.sorted_slots), ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp),
                   ptr(self.stamp), 0, s, n_kernels=7)
        self._adam_untouched(ws)

    def _adam_untouched(self, ws: _Ws, late: bool = False):
        """First half of the split exact-dense update (armed by fused_step only): every large-table row without the fresh stamp."""
        if not late:
            ws.adam_untouched_done = False
        if not ((self._arm_split or late) and self.split_dense_adam and self.adam_mode == TT_ADAM_DENSE_EXACT):
            return
        if self.split_late != late:
            return
        self._call("tt_ebc_backward_adam", ptr(self.arena), ptr(self.arena_m), ptr(self.arena_v), 0, self.total_rows, self.cfg.emb_dim,
                   0, 0, 0, ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp), 0, ptr(self.stamp), 0, 0, 0, 1, 0, 0, 0, 0, 1, 0,
                   self.act, ptr(ws.piece_ws), ws.piece_ws.numel(), ptr(self.hyper), TT_ADAM_DENSE_UNTOUCHED, 0, 0, 0, 0, 0,
                   self._stream())
        ws.adam_untouched_done = True

    def embedding_backward(self, ws: _Ws, apply_adam: bool, want_dense_grad: bool, sorted_already: bool = False):
        """Large tables: (sort-by-row dedup, then) segment-sum of the pooled gradients (+ fused Adam)."""
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        n0, n1 = B * self.KL[0], B * self.KL[1]
        n = n0 + n1
        if n == 0 or self.total_rows == 0:
            return
        H0 = cfg.hidden[0]
        if not sorted_already:
            self.embedding_sort(ws)
        self.profile_tag = ""
        if self.lazy:
            if want_dense_grad or not apply_adam:
                raise native.TwoTowerNativeError("optimizer_mode='lazy_exact' updates the tables inside the library (attach a FusedAdam): "
                                                 "dense table gradients for an external optimizer need optimizer_mode='dense_exact'")
            self._call("tt_ebc_backward_adam_lazy", ptr(self.arena), ptr(self.arena_m), ptr(self.arena_v), self.total_rows, E,
                       ptr(ws.sorted_rows), ptr(ws.sorted_slots), n, ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp),
                       ptr(self.stamp), ptr(ws.dc[0]), H0 + self.K[0] * E, H0, max(self.KL[0], 1), n0, ptr(ws.dc[1]), H0 + self.K[1] * E,
                       H0, max(self.KL[1], 1), n1, self.act, ptr(ws.piece_ws), ws.piece_ws.numel(), ptr(self.hyper), 0, 0, 0,
                       ptr(self.colmap[0]), ptr(self.colmap[1]), ptr(self.lazy_ctx), s, n_kernels=3)
            return
        if want_dense_grad and self.arena_grad is None:
            self.arena_grad = torch.zeros_like(self.arena)
        self._call("tt_ebc_backward_adam", ptr(self.arena) if apply_adam else 0, ptr(self.arena_m) if apply_adam else 0,
                   ptr(self.arena_v) if apply_adam else 0, ptr(self.arena_grad) if want_dense_grad else 0,
                   self.total_rows, E, ptr(ws.sorted_rows), ptr(ws.sorted_slots), n, ptr(self.row_begin), ptr(self.row_end),
                   ptr(self.row_stamp), 0, ptr(self.stamp), ptr(ws.dc[0]), H0 + self.K[0] * E, H0, max(self.KL[0], 1), n0,
                   ptr(ws.dc[1]), H0 + self.K[1] * E, H0, max(self.KL[1], 1), n1, self.act, ptr(ws.piece_ws), ws.piece_ws.numel(),
                   ptr(self.hyper), self._table_mode(ws, apply_adam), 0, 0, 0, ptr(self.colmap[0]),
                   ptr(self.colmap[1]), s, n_kernels=2)

    def _table_mode(self, ws: _Ws, apply_adam: bool) -> int:
        if not apply_adam:
            return TT_ADAM_DENSE_EXACT
        if getattr(ws, "adam_untouched_done", False):  # the untouched rows were updated behind the sort: only the touched ones now
            ws.adam_untouched_done = False
            return TT_ADAM_SPARSE
        return self.adam_mode

    def backward(self, ws: _Ws, dloss: float = 1.0, de_given: bool = False):
        """Backward of the loss and both towers. `de_given`: ws.de already holds the gradients of the two tower outputs
        (a loss evaluated outside the library, e.g. loss_type="cosine_embedding"); only the towers run."""
        if not de_given:
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
        if self.lazy:
            self.note_step()
            self._call("tt_hyper_advance_hist", ptr(self.hyper), lr, warmup_steps, betas[0], betas[1], eps, weight_decay,
                       self.rng_seed, ptr(self.lazy_hist), self.hist_len, self._stream())
            return
        self._call("tt_hyper_advance", ptr(self.hyper), lr, warmup_steps, betas[0], betas[1], eps, weight_decay,
                   self.rng_seed, self._stream())

    # ---- lazy exact-replay Adam bookkeeping (host side)
    def note_step(self):
        """One more optimizer step is about to run (eagerly, or as a graph replay): flush before the history ring could wrap."""
        if not self.lazy:
            return
        if self._steps_since_flush >= self.hist_len - 4 and not torch.cuda.is_current_stream_capturing():
            self.flush_lazy()
        self._steps_since_flush += 1

    def flush_lazy(self):
        """Bring every row of the sorted tables up to the current optimizer step (before weights / moments are read from outside
        the engine: state_dict, checkpoints, gathers). No-op unless optimizer_mode='lazy_exact'."""
        if self.lazy and self.total_rows > 0:
            self._call("tt_ebc_lazy_flush", ptr(self.lazy_ctx), self.total_rows, self.cfg.emb_dim, self._stream())
            self._steps_since_flush = 0

    def lazy_mark_current(self):
        """The tables were just written from outside (load_state_dict / resume): every row is current as of the present step."""
        if self.lazy:
            self.row_last.fill_(int(self.hyper.view(torch.int64)[0].item()))
            self._steps_since_flush = 0

    def dense_adam(self):
        self._call("tt_dense_adam", ptr(self.dense_w), ptr(self.dense_g), ptr(self.dense_m), ptr(self.dense_v),
                   self.P_tot, ptr(self.hyper), self._stream())

    def fused_step(self, notice_dense, notice_values, company_dense, company_values, lr=1e-3, warmup_steps=0,
                   betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, layout=None, masks=None):
        """hyper_advance -> forward -> backward -> sparse-dedup + Adam on tables -> Adam on dense. Returns out5."""
        if self._crit_stream is not None and self.two_streams and self.profile is None:
            # the step's critical chain (lookup -> towers -> logits passes -> tower backward) on a HIGH-priority stream: the caller's
            # stream has the lowest priority there is, i.e. the same as the sort-dedup / weight-gradient helper streams whose CTAs
            # then compete with the persistent logits kernels for SM slots
            outer = torch.cuda.current_stream(self.device)
            self._crit_stream.wait_stream(outer)
            with torch.cuda.stream(self._crit_stream):
                out = self._fused_step(notice_dense, notice_values, company_dense, company_values, lr, warmup_steps, betas, eps,
                                       weight_decay, layout, masks)
            outer.wait_stream(self._crit_stream)
            return out
        return self._fused_step(notice_dense, notice_values, company_dense, company_values, lr, warmup_steps, betas, eps, weight_decay,
                                layout, masks)

    def _fused_step(self, notice_dense, notice_values, company_dense, company_values, lr, warmup_steps, betas, eps, weight_decay,
                    layout, masks):
        main = torch.cuda.current_stream(self.device)
        self._hyper_done = None
        if self.two_streams and self.profile is None and self.hyper_on_helper:
            # nothing before the BatchNorm/dropout kernels reads the optimizer scalars: advance them on a helper stream so the
            # lookups are the first nodes of the step
            hs = self._wg_streams[0][1]
            hs.wait_stream(main)
            with torch.cuda.stream(hs):
                self.hyper_advance(lr, warmup_steps, betas, eps, weight_decay)
                self._hyper_done = torch.cuda.Event()
                self._hyper_done.record(hs)
        else:
            self.hyper_advance(lr, warmup_steps, betas, eps, weight_decay)
        self._arm_split = self._in_fused = True
        self._lazy_offset = -1
        try:
            out5, ws = self.forward(notice_dense, notice_values, company_dense, company_values, True, layout, masks, sort_on_side=True)
        finally:
            self._arm_split = self._in_fused = False
            self._lazy_offset = 0
            self._hyper_done = None
        if self.two_streams:
            # Three streams below the loss: the towers' backward passes on `main` / `_side`, and on `_sort_stream` (where the
            # sort-dedup already ran underneath the forward pass) the large-table segment-sum + Adam, which starts as soon as
            # both concat-buffer gradients exist and overlaps the first-layer weight gradients and the small-table reductions.
            self.loss_backward(ws)
            if self.split_late and self.total_rows > 0:
                self._sort_stream.wait_stream(main)
                with torch.cuda.stream(self._sort_stream):
                    self._adam_untouched(ws, late=True)
            dm = getattr(ws, "deferred_metrics", None)
            # TT_METRICS_LATE=1 launches them behind both towers' data-gradient chains instead: the tower kernels get faster (head
            # 16.6 -> 12.8 us) but the metrics kernel then sits next to the table update at the end of the step and the join waits
            # for it: 315.2 us per step against 295.1 us — off
            late_metrics = dm is not None and os.environ.get("TT_METRICS_LATE", "0") != "0"
            if dm is not None and not late_metrics:
                # Behind the logits passes: the persistent logits kernels need every SM, and small kernels of a side stream launched
                # next to them were measured to delay their CTAs (369 vs 362 us per step).
                dm()
                ws.deferred_metrics = None
            self._side.wait_stream(main)
            self._hi0.wait_stream(main)
            def table_update():  # on `_side`, right behind the company tower's last data gradient
                self._side.wait_event(ws.dc_ready[0])
                self._side.wait_stream(self._sort_stream)
                self._before_table_update(ws)
                self.embedding_backward(ws, apply_adam=True, want_dense_grad=False, sorted_already=True)

            with torch.cuda.stream(self._hi0):
                self.tower_backward(0, ws)
            with torch.cuda.stream(self._side):
                self.tower_backward(1, ws, after_dc=table_update)
            if late_metrics:
                # ... and the loss value / metrics kernels only behind BOTH towers' data-gradient chains (they were measured to slow
                # the fused tower backward kernels by half when launched next to them): they share the machine with the weight
                # gradients and the table update instead, and are joined at the end of the step.
                main.wait_event(ws.dc_ready[0])
                main.wait_event(ws.dc_ready[1])
                dm()
                ws.deferred_metrics = None
            main.wait_stream(self._hi0)
            for ev in getattr(ws, "wg_joins", None) or [ws.dc_ready[1]]:
                main.wait_event(ev)  # the company tower's weight gradients and small-table reductions (helper streams)
            ws.wg_joins = None
            self.reduce_dense_grads()
            self.dense_adam()
            main.wait_stream(self._side)  # ... and the large-table update, which ran next to them
            self._end_of_step(ws)
        else:
            self.backward(ws)
            self.reduce_dense_grads()
            main.wait_stream(self._sort_stream)  # the sort-dedup ran underneath the towers and the loss
            self.embedding_backward(ws, apply_adam=True, want_dense_grad=False, sorted_already=True)
            self.dense_adam()
        return out5

    def reduce_dense_grads(self):
        """Single GPU: nothing. Distributed engines all-reduce the flat dense gradient buffer here."""
        return None

    def _before_table_update(self, ws: _Ws):
        """Hook on the table-update stream once this rank's concat-buffer gradients are complete. Distributed engines put the
        'every rank's pooled gradients are complete' barrier here (the update pulls them over NVLink)."""
        return None

    def _end_of_step(self, ws: _Ws):
        """The very end of the fused step (main stream): join the deferred loss / metrics stream."""
        ev = getattr(ws, "metrics_done", None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            ws.metrics_done = None

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
