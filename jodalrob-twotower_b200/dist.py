"""Multi-GPU two-tower step: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) for the
plumbing, our own kernels over peer-mapped memory for the embedding exchange.

The reference is single-GPU (no torch.distributed anywhere); the oracle for everything here is the
single-device step on the concatenated global batch.

Per step and rank (B_l local samples, B_g = world * B_l):
  ids        clamp -> NCCL all-gather of int32 ids (static size; also the barrier that separates steps)
  lookup     tt_shard_select_owned (deterministic compaction of the ids this rank owns) -> tt_ebc_push_rows: rows are
             gathered and stored directly into the requesters' concat buffers through peer pointers (fused lookup +
             all-to-all) -> light barrier
  towers     replicated dense weights; BatchNorm statistics all-reduced (SyncBN == BN on the global batch)
  loss       all-gather both towers' embeddings; each rank owns its rows of S: row statistics are local, column
             statistics are all-reduced; global diagonal = rank * B_l + i; loss means are over B_g
  backward   dN from (local notice rows x all company columns), dC from (local company rows x all notice columns) —
             the transposed problem, so no reduce-scatter is needed; dense grads all-reduced (flat buffer);
  tables     sort-dedup of the owned ids, then the segment-sum kernel PULLS the pooled gradients out of the
             requesters' buffers over NVLink (fused gradient all-to-all) and applies Adam to the local shard.
"""
from __future__ import annotations

import os

import math
import threading
from typing import Dict, List, Optional, Sequence

import torch

from . import native
from .engine import TwoTowerConfig, TwoTowerEngine, _Ws
from .native import TT_ADAM_DENSE_EXACT, TT_F32, ptr


# ------------------------------------------------------------------------------------------------ shard plan (pure integers)
class ShardPlan:
    """Row-wise shard map. block_k = ceil(V_k / world); owner = id // block_k; local = id - owner * block_k."""

    def __init__(self, vocab: Sequence[int], world: int):
        self.world = int(world)
        self.vocab = [int(v) for v in vocab]
        self.block = [-(-v // self.world) for v in self.vocab]
        self.local_base = [0]
        for b in self.block[:-1]:
            self.local_base.append(self.local_base[-1] + b)
        self.local_rows = sum(self.block)

    def owner(self, k: int, idx: int) -> int:
        return idx // self.block[k]

    def local_row(self, k: int, idx: int) -> int:
        return self.local_base[k] + idx - self.owner(k, idx) * self.block[k]

    def shard_range(self, k: int, rank: int):
        """global id range [lo, hi) of table k held by `rank` (may be empty for the last ranks of tiny tables)."""
        lo = min(rank * self.block[k], self.vocab[k])
        return lo, min(lo + self.block[k], self.vocab[k])


# ------------------------------------------------------------------------------------------------ communicators
class NcclComm:
    """torch.distributed backend (NCCL). Peer-mapped buffers come from torch symmetric memory."""
    peer_kernels = True  # one process per GPU: kernels of different ranks may wait for each other (tt_peer_allreduce_f64)

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self._tiny = None

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        self.dist.all_gather_into_tensor(out, inp, group=self.group)

    def all_reduce(self, t: torch.Tensor):
        self.dist.all_reduce(t, group=self.group)

    def barrier(self, device):
        if self._tiny is None:
            self._tiny = torch.zeros(1, dtype=torch.float32, device=device)
        self.dist.all_reduce(self._tiny, group=self.group)

    def clone(self) -> "NcclComm":
        """A second communicator over the same ranks: collectives issued on another stream (the company tower) must not share
        a communicator with the main stream's, NCCL orders the operations of one communicator."""
        return NcclComm(self.dist.new_group(ranks=self.dist.get_process_group_ranks(self.group), backend="nccl"))

    def peer_buffer(self, nbytes: int, device) -> "tuple[torch.Tensor, List[int]]":
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(t, self.group)
        self._keep = getattr(self, "_keep", []) + [hdl]
        return t, [int(p) for p in hdl.buffer_ptrs]


class ThreadWorld:
    """In-process emulation of `world` ranks on ONE GPU, one Python thread per rank; collectives are host-mediated
    (no kernel ever waits on another), so it is safe on a single device. Used by the -m gpu tests to cover the N>1 path."""

    def __init__(self, world: int):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots: List[Optional[torch.Tensor]] = [None] * world
        self.ptrs: List[Optional[int]] = [None] * world


class ThreadComm:
    peer_kernels = False  # all "ranks" share one GPU: collectives stay host-mediated, no kernel ever waits on another

    def __init__(self, tw: ThreadWorld, rank: int):
        self.tw, self.world, self.rank = tw, tw.world, rank

    def _exchange(self, t: torch.Tensor) -> List[torch.Tensor]:
        torch.cuda.current_stream(t.device).synchronize()
        self.tw.slots[self.rank] = t
        self.tw.bar.wait()
        got = list(self.tw.slots)
        return got

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        got = self._exchange(inp)
        n = inp.numel()
        flat = out.view(-1)
        for r, t in enumerate(got):
            flat[r * n:(r + 1) * n].copy_(t.view(-1))
        torch.cuda.current_stream(out.device).synchronize()
        self.tw.bar.wait()

    def all_reduce(self, t: torch.Tensor):
        got = self._exchange(t)
        acc = got[0].clone()
        for o in got[1:]:
            acc += o  # fixed rank order on every rank -> identical result everywhere
        torch.cuda.current_stream(t.device).synchronize()
        self.tw.bar.wait()
        t.copy_(acc)
        torch.cuda.current_stream(t.device).synchronize()
        self.tw.bar.wait()

    def barrier(self, device):
        torch.cuda.current_stream(device).synchronize()
        self.tw.bar.wait()

    def clone(self):
        return self  # host-mediated and blocking: one rank thread never has two collectives in flight

    def peer_buffer(self, nbytes: int, device):
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.tw.ptrs[self.rank] = t.data_ptr()
        self.tw.bar.wait()
        ptrs = [int(p) for p in self.tw.ptrs]
        self.tw.bar.wait()
        return t, ptrs


class SingleComm:
    """world == 1: the sharded pipeline with itself as the only peer (exercises select/push/pull on one GPU)."""
    world, rank = 1, 0
    peer_kernels = False

    def all_gather(self, out, inp):
        out.view(-1).copy_(inp.view(-1))

    def all_reduce(self, t):
        return None

    def barrier(self, device):
        return None

    def clone(self):
        return self

    def peer_buffer(self, nbytes, device):
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        return t, [t.data_ptr()]


# ------------------------------------------------------------------------------------------------ engine
def _a256(n: int) -> int:
    return (n + 255) // 256 * 256


class DistributedEngine(TwoTowerEngine):
    """Small tables (V*E <= SMALL_TABLE_FLOATS) are replicated with the dense weights and their gradients ride the dense
    all-reduce; large tables are sharded row-wise and exchanged over peer memory."""

    def __init__(self, cfg: TwoTowerConfig, device, comm, optimizer_mode: str = "dense_exact", value_layout: str = "sample_major",
                 dtype: str = "fp32", small_tables: bool = True, bag_len: int = 1):
        self.comm = comm
        self.world, self.rank = comm.world, comm.rank
        self.full_cfg = cfg
        # Multi-hot bags (BASELINE configs[4]): every (sample, key) bag holds exactly `bag_len` ids, values [B, K, L] flattened
        # sample-major; sum pooling (TorchRec EBC semantics; the reference's bags hold one id, unified_bid_data_loader.py:835).
        # Owners pre-pool their ids of a bag; all tables go through the sharded path.
        if bag_len > 1:
            if optimizer_mode == "lazy_exact" or value_layout != "sample_major":
                raise ValueError("bag_len > 1 supports optimizer_mode 'dense_exact' / 'sparse' and the sample-major value layout")
            small_tables = False
        super().__init__(cfg, device, optimizer_mode, value_layout, dtype, small_tables)
        self.bag_len = int(bag_len)
        dev = self.device
        # large keys in slot order (notice large keys, then company large keys) and their shard map
        self.large_keys = [[k for k in range(self.K[t]) if not self.is_small[t][k]] for t in range(2)]
        lv = [cfg.tower(t)[2][k] for t in range(2) for k in self.large_keys[t]]
        self.plan = ShardPlan(lv if lv else [1], self.world)
        L = self.bag_len
        rep = lambda xs: [x for x in xs for _ in range(L)]
        # routing sees one pseudo-key per (key, position in the bag): slot = b * (KL * L) + kl * L + l
        self.block = torch.tensor(rep(self.plan.block), dtype=torch.int32, device=dev)
        self.local_base = torch.tensor(rep(self.plan.local_base), dtype=torch.int32, device=dev)
        self.d_large_keys = [torch.tensor([k * L + l for k in ks for l in range(L)] if ks else [0], dtype=torch.int32, device=dev)
                             for ks in self.large_keys]
        self.KLp = (self.KL[0] * L, self.KL[1] * L)
        self.vocab_p = torch.tensor(rep(self.vocab.tolist()), dtype=torch.int32, device=dev)
        self.colmap_p = [torch.tensor(rep(c.tolist()), dtype=torch.int32, device=dev) for c in self.colmap]
        # the local mixed lookup serves the replicated small tables only: large keys are marked "skip" (vocab 0)
        vf = self.vocab.clone()
        off = 0
        for t in range(2):
            for k in self.large_keys[t]:
                vf[off + k] = 0
            off += self.K[t]
        self.vocab_small_only = vf
        self.no_rank = torch.full((self.K[0] + self.K[1],), -1, dtype=torch.int32, device=dev)
        self.rng_stream_base = 64 * self.rank
        self.bn_affine_grad_scale = 1.0 if self.rank == 0 else 0.0  # the SyncBN sums are global: one rank publishes them
        # the company tower runs on the side stream with its own communicator for the SyncBN all-reduces
        self.comm_t = [comm, comm.clone()]
        self._setup_peer_sites()
        self.two_streams = True
        self.sharded_lookup = True
        if lv:
            assert self.total_rows == self.plan.local_rows, (self.total_rows, self.plan.local_rows)

    # ---- tiny reductions over NVLink peer memory instead of collective-library calls (csrc/peer.cu)
    PEER_SITES = ("tower0", "tower1", "partial", "barrier", "ids", "e0", "e1", "colsum", "invrow", "dcbar")

    def _setup_peer_sites(self):
        self.peer_sites = None
        if not getattr(self.comm, "peer_kernels", False) or self.world == 1:
            return
        dev = self.device
        if "TT_PEER_TIMEOUT_MS" in os.environ:  # watchdog of the peer-wait kernels (default 10 min; 0 = wait for ever)
            self.lib.call("tt_peer_set_timeout_ms", int(os.environ["TT_PEER_TIMEOUT_MS"]))
        self.peer_n_max = max(2 * max(self.cfg.hidden[1:] + [1]), 128)
        mb = self.lib.query("tt_peer_mailbox_bytes", self.world, self.peer_n_max)
        mb = (mb + 255) // 256 * 256
        buf, peers = self.comm.peer_buffer(mb * len(self.PEER_SITES), dev)
        buf.zero_()
        torch.cuda.synchronize(dev)
        self.comm.dist.barrier(group=self.comm.group, device_ids=[dev.index if dev.index is not None else torch.cuda.current_device()])  # every mailbox is zero before anyone writes
        self._peer_buf = buf
        self.peer_sites = {name: torch.tensor([p + i * mb for p in peers], dtype=torch.int64, device=dev)
                           for i, name in enumerate(self.PEER_SITES)}
        self.peer_epochs = torch.zeros(len(self.PEER_SITES), dtype=torch.int64, device=dev)
        self.peer_counts = torch.zeros(len(self.PEER_SITES), dtype=torch.int32, device=dev)
        self._small_stream = torch.cuda.Stream(dev)
        self._peer_dummy = torch.zeros(1, dtype=torch.float64, device=dev)
        self._peer_dummy2 = torch.zeros(1, dtype=torch.float64, device=dev)
        self._peer_dummy3 = torch.zeros(1, dtype=torch.float64, device=dev)

    def _peer_allreduce(self, buf: torch.Tensor, n: int, site: str):
        i = self.PEER_SITES.index(site)
        self._call("tt_peer_allreduce_f64", ptr(buf), n, self.rank, self.world, ptr(self.peer_sites[site]), self.peer_n_max,
                   self.peer_epochs.data_ptr() + 8 * i, self._stream())

    def _peer_allgather(self, src: torch.Tensor, dst_table: torch.Tensor, site: str):
        """src (this rank's contiguous block) -> block `rank` of every rank's destination buffer; returns (on the stream) when all
        W blocks have landed here (csrc/peer.cu)."""
        i = self.PEER_SITES.index(site)
        nbytes = src.numel() * src.element_size()
        self._call("tt_peer_allgather", ptr(src), nbytes, self.rank, self.world, ptr(dst_table), self.rank * nbytes,
                   ptr(self.peer_sites[site]), self.peer_epochs.data_ptr() + 8 * i, self.peer_counts.data_ptr() + 4 * i, self._stream())

    def _alloc_rows(self, t: int, k: int, vocab: int, small: bool) -> int:
        return vocab if small else -(-vocab // self.world)

    # ---- workspace: concat / gradient buffers live in one peer-mapped allocation
    def workspace(self, B: int) -> _Ws:
        ws = self._ws.get(B)
        if ws is not None:
            return ws
        cfg, dev, E, W = self.cfg, self.device, self.cfg.emb_dim, self.world
        H0 = cfg.hidden[0]
        ld = [H0 + self.K[t] * E for t in range(2)]
        esz = 4 if self.act == TT_F32 else 2
        sizes = [_a256(B * ld[t] * esz) for t in range(2)]
        total = 2 * (sizes[0] + sizes[1])
        buf, peer = self.comm.peer_buffer(total, dev)
        offs = [0, sizes[0], sizes[0] + sizes[1], 2 * sizes[0] + sizes[1]]
        views = [buf[offs[i]:offs[i] + B * ld[i % 2] * esz].view(self.act_torch).view(B, ld[i % 2]) for i in range(4)]
        # build the generic workspace first, then swap the four buffers for the peer-mapped ones
        ws = super().workspace(B)
        ws.c, ws.dc = [views[0], views[1]], [views[2], views[3]]
        ws.peer_buf = buf
        ws.c_table = torch.tensor([p + offs[t] for p in peer for t in range(2)], dtype=torch.int64, device=dev)
        ws.dc_table = torch.tensor([p + offs[2 + t] for p in peer for t in range(2)], dtype=torch.int64, device=dev)
        n_l = B * (self.KLp[0] + self.KLp[1])
        cap = max(n_l * W, 1)
        i32 = dict(dtype=torch.int32, device=dev)
        ws.ids_local = torch.zeros(max(n_l + 3, 4) // 4 * 4, **i32)[:max(n_l, 1)]
        ws.ids_all = torch.zeros(cap, **i32)
        ws.rows = torch.zeros(cap, **i32)          # owned local rows (compact)
        ws.slots = torch.zeros(cap, **i32)         # their global slots
        ws.sorted_rows = torch.zeros(cap, **i32)
        ws.sorted_slots = torch.zeros(cap, **i32)
        ws.n_owned = torch.zeros(1, **i32)
        q = self.lib.query
        ws.shard_ws = torch.empty(q("tt_shard_workspace_bytes", cap), dtype=torch.uint8, device=dev)
        # all-to-all id routing: receive buffers [W senders][n_l] for (local row, global slot) + the W segment lengths, peer-mapped
        ws.part = None
        if self.bag_len > 1:  # per-(owner, bag) partial sums land here (peer-mapped): f32 [W][n_bags][E]
            n_bags = B * (self.KL[0] + self.KL[1])
            pbuf, ppeer = self.comm.peer_buffer(W * n_bags * E * 4, dev)
            ws.part = pbuf
            ws.part_table = torch.tensor(ppeer, dtype=torch.int64, device=dev)
        ws.a2a = None
        if n_l > 0 and (os.environ.get("TT_A2A", "1") != "0" or self.bag_len > 1):
            seg = _a256(cap * 4)
            abuf, apeer = self.comm.peer_buffer(2 * seg + 256, dev)
            # NOT zeroed here: every segment length is written by its sender before the owner reads it, and a local memset with no
            # cross-rank barrier behind it could wipe what a faster peer has already pushed for the first step
            ws.a2a_buf = abuf
            ws.a2a = dict(rows=abuf[:cap * 4].view(torch.int32), slots=abuf[seg:seg + cap * 4].view(torch.int32),
                          counts=abuf[2 * seg:2 * seg + 4 * W].view(torch.int32),
                          t_rows=torch.tensor(apeer, dtype=torch.int64, device=dev),
                          t_slots=torch.tensor([p_ + seg for p_ in apeer], dtype=torch.int64, device=dev),
                          t_counts=torch.tensor([p_ + 2 * seg for p_ in apeer], dtype=torch.int64, device=dev),
                          ws=torch.empty(q("tt_shard_a2a_workspace_bytes", max(n_l, 1), W), dtype=torch.uint8, device=dev))
        ws.sort_ws = torch.empty(q("tt_ebc_sort_workspace_bytes", cap, max(self.total_rows, 1)), dtype=torch.uint8, device=dev)
        ws.piece_ws = torch.empty(q("tt_ebc_piece_workspace_bytes", cap, E), dtype=torch.uint8, device=dev)
        Bg = B * W
        f32 = dict(dtype=torch.float32, device=dev)
        ws.e_all = [torch.empty(Bg, cfg.final_dim, dtype=self.emb_torch, device=dev) for _ in range(2)]
        ws.col_sumexp = torch.empty(Bg, **f32)
        ws.col_sum_s = torch.empty(Bg, **f32)
        ws.inv_col = torch.empty(Bg, **f32)
        ws.inv_row_all = torch.empty(Bg, **f32)
        ws.xtab = None
        # Exchange destinations in peer-mapped memory: with peer kernels available the id exchange, the embedding all-gathers,
        # the column-sum all-reduce and the row-scale all-gather are pushes over NVLink (tt_peer_allgather) instead of NCCL calls.
        # Every block must be a multiple of 16 bytes.
        esz_e = 4 if self.emb_tt == TT_F32 else 2
        blocks_ok = (n_l * 4) % 16 == 0 and (B * cfg.final_dim * esz_e) % 16 == 0 and (B * 4) % 16 == 0 and n_l > 0
        if self.peer_sites is not None and blocks_ok and os.environ.get("TT_PEER_EXCHANGE", "1") != "0":
            parts = {"ids": cap * 4, "e0": Bg * cfg.final_dim * esz_e, "e1": Bg * cfg.final_dim * esz_e,
                     "colsum": W * (Bg + B) * 4,  # per rank: its column-sum partials [Bg] followed by its row sums [B]
                     "rows3": Bg * 4, "cols3": Bg * 4}  # three-pass step: all ranks' row sums / column sums of exp
            xoff, o = {}, 0
            for k_, nb in parts.items():
                xoff[k_] = o
                o += _a256(nb)
            xbuf, xpeer = self.comm.peer_buffer(o, dev)
            xbuf.zero_()
            ws.xbuf = xbuf
            ws.ids_all = xbuf[xoff["ids"]:xoff["ids"] + cap * 4].view(torch.int32)
            ws.e_all = [xbuf[xoff[k_]:xoff[k_] + parts[k_]].view(self.emb_torch).view(Bg, cfg.final_dim) for k_ in ("e0", "e1")]
            ws.colsum_slots = xbuf[xoff["colsum"]:xoff["colsum"] + W * (Bg + B) * 4].view(torch.float32)
            ws.x_rows3 = xbuf[xoff["rows3"]:xoff["rows3"] + Bg * 4].view(torch.float32)
            ws.x_cols3 = xbuf[xoff["cols3"]:xoff["cols3"] + Bg * 4].view(torch.float32)
            ws.xmsg = torch.empty(Bg + B, **f32)      # what this rank sends: column-sum partials, then its row sums
            ws.col_partial, ws.row_sumexp = ws.xmsg[:Bg], ws.xmsg[Bg:]
            ws.xtab = {k_: torch.tensor([p_ + xoff[k_] for p_ in xpeer], dtype=torch.int64, device=dev) for k_ in parts}
            torch.cuda.synchronize(dev)
            self.comm.dist.barrier(group=self.comm.group, device_ids=[dev.index if dev.index is not None else torch.cuda.current_device()])
        # the four-pass entry points run this rank's B rows against all Bg columns; the size is NOT monotone in the row count (a
        # short, wide problem splits a row block over more CTAs: more partials per row) — at W = 8 and B = 256 the square size was
        # too small ("workspace too small" in the fp32 check of tests/dist_check.py on 8 GPUs)
        ws.sim_ws = torch.empty(max(q("tt_simce_workspace_bytes", B, Bg, cfg.final_dim), q("tt_simce_workspace_bytes", Bg, Bg, cfg.final_dim)),
                                dtype=torch.uint8, device=dev)
        # three-pass logits step: this rank's rows against all ranks' columns; the row / column sums of the other ranks arrive
        # by all-gather (peer pushes when available)
        self._alloc_sim3(ws, B, Bg)
        ws.rowsum_all = torch.empty(Bg, **f32)
        ws.colsum_all = torch.empty(Bg, **f32)
        ws.sums_loc = [torch.empty(B, **f32) for _ in range(2)]
        if ws.xtab is not None and "rows3" in ws.xtab:
            ws.rowsum_all, ws.colsum_all = ws.x_rows3, ws.x_cols3
        return ws

    # ---- forward pieces
    def embedding_forward(self, ws: _Ws, values, layout: int):
        E, B, s = self.cfg.emb_dim, ws.B, self._stream()
        H0 = self.cfg.hidden[0]
        K0, K1 = self.K
        KL0, KL1 = self.KLp  # (key, position) pseudo-keys: == the large keys when bags hold one id
        L = self.bag_len
        n0, n_l = B * KL0, B * (KL0 + KL1)
        if values[0] is None or values[1] is None:
            raise NotImplementedError("the sharded lookup serves both towers in one exchange; call the train task, not a single tower")
        self.profile_tag = "sharded"
        # replicated small tables: plain local lookup (large keys skipped) — on a stream of its own when the exchange below can
        # run next to it (it writes other columns of the concat buffers)
        cur = torch.cuda.current_stream(self.device)
        par = self.peer_sites is not None and self.profile is None and n_l > 0
        small_stream = self._small_stream if par else cur
        if par:
            small_stream.wait_stream(cur)
        with torch.cuda.stream(small_stream):
            for t, v in enumerate(values):
                k0 = 0 if t == 0 else K0
                if self.n_small[t]:
                    self._call("tt_ebc_forward_mixed", ptr(self.flat_w), self.row_base.data_ptr() + 8 * k0,
                               self.vocab_small_only.data_ptr() + 4 * k0, ptr(v), B, self.K[t], E, layout, ptr(ws.c[t]), self.act,
                               H0 + self.K[t] * E, H0, self.no_rank.data_ptr() + 4 * k0, 0, 0, 0, small_stream.cuda_stream)
        ws.values, ws.layout = list(values), layout
        if n_l == 0:
            return
        # sharded large tables: clamp -> all-gather ids -> owned-id compaction -> fused lookup + push over peer memory
        if ws.part is not None:  # owners only write the (owner, bag) partials they have: everything else must read as zero
            self._call("tt_zero_bytes", ptr(ws.part), ws.part.numel(), s)
        self._call("tt_ebc_clamp_ids_subset", ptr(values[0]), ptr(self.vocab_p), ptr(self.d_large_keys[0]), B, K0 * L, KL0, layout,
                   ptr(ws.ids_local), s)
        self._call("tt_ebc_clamp_ids_subset", ptr(values[1]), self.vocab_p.data_ptr() + 4 * K0 * L, ptr(self.d_large_keys[1]), B, K1 * L,
                   KL1, layout, ws.ids_local.data_ptr() + 4 * n0, s)
        if ws.a2a is not None:
            # all-to-all: bucket my ids by owner, push (local row, slot) into the owners' receive buffers over NVLink, barrier, pack
            a = ws.a2a
            self._call("tt_shard_a2a_send", ptr(ws.ids_local), ptr(self.block), ptr(self.local_base), n_l, n0, max(KL0, 1), KL1,
                       self.world, self.rank, ptr(a["t_rows"]), ptr(a["t_slots"]), ptr(a["t_counts"]), ptr(a["ws"]), a["ws"].numel(), s,
                       n_kernels=3)
            if self.peer_sites is not None:
                self._peer_allreduce(self._peer_dummy3, 1, "ids")
            else:
                self.comm.barrier(self.device)
            self._call("tt_shard_a2a_compact", ptr(a["rows"]), ptr(a["slots"]), ptr(a["counts"]), self.world, n_l, ptr(ws.rows),
                       ptr(ws.slots), ptr(ws.n_owned), s)
        else:
            if ws.xtab is not None:
                self._peer_allgather(ws.ids_local, ws.xtab["ids"], "ids")
            else:
                self.comm.all_gather(ws.ids_all, ws.ids_local)
            self._call("tt_shard_select_owned", ptr(ws.ids_all), ptr(self.block), ptr(self.local_base), n_l, n0, max(KL0, 1), KL1,
                       self.world, self.rank, ptr(ws.rows), ptr(ws.slots), ptr(ws.n_owned), ptr(ws.shard_ws), ws.shard_ws.numel(), s,
                       n_kernels=3)
        if L > 1:
            # owners pre-pool their ids of every bag and store one partial per (owner, bag) at the requester; after the barrier the
            # requester folds the W partials in rank order into its concat buffers
            n_bags = B * (self.KL[0] + self.KL[1])
            self._call("tt_ebc_push_pooled", ptr(self.arena), E, ptr(ws.rows), ptr(ws.slots), ptr(ws.n_owned), n_l * self.world,
                       ptr(ws.part_table), n_l, L, n_bags, self.rank, s)
            if self.peer_sites is not None:
                self._peer_allreduce(self._peer_dummy, 1, "barrier")
            else:
                self.comm.barrier(self.device)
            self._call("tt_ebc_sum_partials", ptr(ws.part), self.world, n_bags, E, ptr(ws.c[0]), H0 + K0 * E, ptr(self.colmap[0]),
                       B * self.KL[0], max(self.KL[0], 1), ptr(ws.c[1]), H0 + K1 * E, ptr(self.colmap[1]), self.KL[1], self.act, s)
            if par:
                cur.wait_stream(small_stream)
            return
        if self.lazy:  # lazy exact-replay Adam: the owner catches a row up (read-only) before it sends it
            self._call("tt_ebc_push_rows_lazy", ptr(self.arena), E, ptr(ws.rows), ptr(ws.slots), ptr(ws.n_owned), n_l * self.world,
                       ptr(ws.c_table), n_l, n0, max(KL0, 1), KL1, H0 + K0 * E, H0 + K1 * E, H0, H0, ptr(self.colmap[0]),
                       ptr(self.colmap[1]), self.act, ptr(self.lazy_ctx), self._lazy_offset, s)
        else:
            self._call("tt_ebc_push_rows", ptr(self.arena), E, ptr(ws.rows), ptr(ws.slots), ptr(ws.n_owned), n_l * self.world,
                       ptr(ws.c_table), n_l, n0, max(KL0, 1), KL1, H0 + K0 * E, H0 + K1 * E, H0, H0, ptr(self.colmap[0]),
                       ptr(self.colmap[1]), self.act, s)
        # every owner has finished writing into my concat buffers
        if self.peer_sites is not None:
            self._peer_allreduce(self._peer_dummy, 1, "barrier")
        else:
            self.comm.barrier(self.device)
        if par:
            cur.wait_stream(small_stream)

    def bn_allreduce(self, sums: torch.Tensor, n: int, local_count: int, tower: int = 0) -> int:
        if self.peer_sites is not None and n <= self.peer_n_max:
            self._peer_allreduce(sums, n, "tower%d" % tower)
        else:
            self.comm_t[tower].all_reduce(sums[:n])
        return local_count * self.world

    def bn_sync(self, part: torch.Tensor, n_tiles: int, n: int, local_count: int, tower: int, sums: torch.Tensor):
        # fused tower kernels: fold this rank's per-tile partials (fp64, tile order), all-reduce, hand the global sums on
        self._call("tt_colpart_reduce_f64", ptr(part), n_tiles, n, ptr(sums), self._stream())
        return self.bn_allreduce(sums, n, local_count, tower), sums

    # fused_step is the base engine's: the large-table update starts on the side stream as soon as this rank's concat-buffer
    # gradients exist AND every other rank's do (barrier below) — it pulls them over NVLink — and overlaps the weight gradients;
    # the all-reduce of the dense gradients runs at the end of the step, next to it.
    def _before_table_update(self, ws: _Ws):
        if self.peer_sites is not None:
            self._peer_allreduce(self._peer_dummy2, 1, "dcbar")
        else:
            self.comm.barrier(self.device)

    def loss_forward(self, ws: _Ws):
        cfg, B, s = self.cfg, ws.B, self._stream()
        Bg, off = B * self.world, B * self.rank
        self.profile_tag = ""
        inv_t = 1.0 / cfg.temperature
        main = torch.cuda.current_stream(self.device)
        px = ws.xtab is not None
        if px:  # the two embedding all-gathers as concurrent pushes over NVLink
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self._peer_allgather(ws.e[1], ws.xtab["e1"], "e1")
            self._peer_allgather(ws.e[0], ws.xtab["e0"], "e0")
            main.wait_stream(self._side)
        else:
            self.comm.all_gather(ws.e_all[0], ws.e[0])
            self.comm.all_gather(ws.e_all[1], ws.e[1])
        css = ptr(ws.col_sum_s) if cfg.label_smoothing > 0 else 0
        ws.pending3 = False
        if self._in_fused and getattr(ws, "sim3_ws", None) is not None:
            # three-pass step (engine.loss_forward): P1 over this rank's notice rows x all company columns, then the row sums of
            # every rank are exchanged and all notice rows scaled by g / row_sum (the second operand of P2)
            F, nb = cfg.final_dim, ws.sim3_ws.numel()
            ws.gscale3 = 0.5 / Bg * inv_t
            self._call("tt_simce3_rows", ptr(ws.es[0]), ptr(ws.e_all[1]), B, Bg, F, inv_t, off, ptr(ws.diag), ptr(ws.sim3_ws), nb, s)
            self._call("tt_simce3_rows_finish", ptr(ws.sim3_ws), nb, B, Bg, F, inv_t, ptr(ws.sums_loc[0]), ptr(ws.inv_row), 0,
                       ws.gscale3, 0, s)
            if px:
                self._peer_allgather(ws.sums_loc[0], ws.xtab["rows3"], "invrow")
            else:
                self.comm.all_gather(ws.rowsum_all, ws.sums_loc[0])
            self._call("tt_simce3_scale_rows", ptr(ws.e_all[0]), ptr(ws.rowsum_all), ws.gscale3, inv_t, Bg, F, ptr(ws.xs[0]), s)
            ws.pending3 = True
            return ws.out5
        self._call("tt_simce_forward", ptr(ws.e[0]), ptr(ws.e_all[1]), B, Bg, cfg.final_dim, self.emb_tt, inv_t, off,
                   ptr(ws.row_sumexp), ptr(ws.col_partial if px else ws.col_sumexp), ptr(ws.row_sum_s), css, ptr(ws.diag),
                   ptr(ws.row_argmax), ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=4)
        if px:
            # ONE exchange for the column-sum all-reduce and the row-scale all-gather: every rank pushes [its column-sum partials |
            # its row sums] into slot `rank` of every rank; a fixed-order sum gives the column sums, a reciprocal the row scales
            self._peer_allgather(ws.xmsg, ws.xtab["colsum"], "colsum")
            self._call("tt_sum_slots_f32", ptr(ws.colsum_slots), self.world, Bg, B, ptr(ws.col_sumexp), ptr(ws.inv_row_all), s)
        else:
            self.comm.all_reduce(ws.col_sumexp)
        if css:
            self.comm.all_reduce(ws.col_sum_s)
        self._call("tt_simce_finalize", ptr(ws.row_sumexp), ptr(ws.col_sumexp), ptr(ws.row_sum_s), css,
                   ptr(ws.diag), ptr(ws.row_argmax), B, Bg, off, Bg, inv_t, cfg.label_smoothing, ptr(ws.inv_row),
                   ptr(ws.inv_col), ptr(ws.partial), s)

        def metrics():
            if self.peer_sites is not None:
                self._peer_allreduce(ws.partial, ws.partial.numel(), "partial")
            else:
                self.comm.all_reduce(ws.partial)
            self._call("tt_simce_scalars", ptr(ws.partial), Bg, ptr(ws.out5), self._stream())

        if px and self._in_fused and self.profile is None:
            # the loss value / metrics are not needed by the backward pass: off the critical path, joined at the end of the step
            self._metrics_stream.wait_stream(main)
            with torch.cuda.stream(self._metrics_stream):
                metrics()
                ws.metrics_done = torch.cuda.Event()
                ws.metrics_done.record(self._metrics_stream)
        else:
            metrics()
        if not px:
            self.comm.all_gather(ws.inv_row_all, ws.inv_row)
        return ws.out5

    def loss_backward(self, ws: _Ws, dloss: float = 1.0):
        cfg, B, s = self.cfg, ws.B, self._stream()
        Bg, off = B * self.world, B * self.rank
        self.profile_tag = ""
        inv_t = 1.0 / cfg.temperature
        gscale = dloss * 0.5 / Bg * inv_t
        F = cfg.final_dim
        if getattr(ws, "pending3", False):
            ws.pending3 = False
            px = ws.xtab is not None
            nb, ls, g = ws.sim3_ws.numel(), cfg.label_smoothing, ws.gscale3
            css = ptr(ws.col_sum_s) if ls > 0 else 0
            # P2: this rank's company rows x all notice columns: their sums of exp ARE the column sums of S for this rank's columns
            # (complete: every notice row is here) — no cross-rank reduction, only an all-gather for P3's operand
            self._call("tt_simce3_grad", ptr(ws.es[1]), ptr(ws.e_all[0]), ptr(ws.xs[0]), B, Bg, F, inv_t, off, 0, g, ls,
                       ptr(ws.sim3_ws), nb, s)
            self._call("tt_simce3_grad_finish", ptr(ws.sim3_ws), nb, B, Bg, F, inv_t, off, ptr(ws.e_all[0]), 1, g, ls, Bg,
                       ptr(ws.de[1]), ptr(ws.sums_loc[1]), ptr(ws.inv_scratch[1]), 0, 0, s)
            if px:
                self._peer_allgather(ws.sums_loc[1], ws.xtab["cols3"], "colsum")
            else:
                self.comm.all_gather(ws.colsum_all, ws.sums_loc[1])
            self._call("tt_simce3_scale_rows", ptr(ws.e_all[1]), ptr(ws.colsum_all), g, inv_t, Bg, F, ptr(ws.xs[1]), s)
            self._call("tt_simce3_grad", ptr(ws.es[0]), ptr(ws.e_all[1]), ptr(ws.xs[1]), B, Bg, F, inv_t, off, ptr(ws.inv_row), g, ls,
                       ptr(ws.sim3_ws), nb, s)
            self._call("tt_simce3_grad_finish", ptr(ws.sim3_ws), nb, B, Bg, F, inv_t, off, ptr(ws.e_all[1]), 0, g, ls, Bg,
                       ptr(ws.de[0]), 0, 0, 0, 0, s)

            def metrics(ms):
                self._call("tt_simce3_metrics", ptr(ws.es[0]), ptr(ws.e_all[1]), B, Bg, F, ptr(ws.row_sum_s), ptr(ws.row_argmax),
                           ptr(ws.sim3_ws), nb, ms, n_kernels=2)
                if css:  # sum_i S_ij / T of this rank's columns: company rows against the sum of ALL notice rows, complete locally
                    self._call("tt_simce3_metrics", ptr(ws.es[1]), ptr(ws.e_all[0]), B, Bg, F, ws.col_sum_s.data_ptr() + 4 * off, 0,
                               ptr(ws.sim3_ws), nb, ms, n_kernels=2)
                self._call("tt_simce_finalize", ptr(ws.sums_loc[0]), ptr(ws.colsum_all), ptr(ws.row_sum_s), css, ptr(ws.diag),
                           ptr(ws.row_argmax), B, Bg, off, Bg, inv_t, ls, ptr(ws.inv_scratch[0]), ptr(ws.inv_col), ptr(ws.partial), ms)
                if self.peer_sites is not None:
                    self._call("tt_peer_allreduce_f64", ptr(ws.partial), ws.partial.numel(), self.rank, self.world,
                               ptr(self.peer_sites["partial"]), self.peer_n_max,
                               self.peer_epochs.data_ptr() + 8 * self.PEER_SITES.index("partial"), ms)
                else:
                    self.comm.all_reduce(ws.partial)
                self._call("tt_simce_scalars", ptr(ws.partial), Bg, ptr(ws.out5), ms)

            if px and self.profile is None and self.defer_metrics:
                def deferred():
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
        self._call("tt_simce_backward", ptr(ws.e[0]), ptr(ws.e_all[1]), B, Bg, F, self.emb_tt, inv_t, off, ptr(ws.inv_row),
                   ptr(ws.inv_col), gscale, cfg.label_smoothing, Bg, ptr(ws.de[0]), ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=2)
        self._call("tt_simce_backward", ptr(ws.e[1]), ptr(ws.e_all[0]), B, Bg, F, self.emb_tt, inv_t, off,
                   ws.inv_col.data_ptr() + 4 * off, ptr(ws.inv_row_all), gscale, cfg.label_smoothing, Bg, ptr(ws.de[1]),
                   ptr(ws.sim_ws), ws.sim_ws.numel(), s, n_kernels=2)

    def reduce_dense_grads(self):
        # dense weights AND the replicated small tables; also the barrier that guarantees every rank's pooled-gradient
        # buffers are complete before anyone pulls from them
        self.comm.all_reduce(self.dense_g)

    def embedding_sort(self, ws: _Ws):
        B, s = ws.B, self._stream()
        n_l = B * (self.KLp[0] + self.KLp[1])
        if n_l == 0 or self.total_rows == 0:
            return
        cap = n_l * self.world
        self.profile_tag = "sharded"
        self._call("tt_ebc_sort_rows_advance", ptr(ws.rows), ptr(ws.slots), cap, self.total_rows, ptr(ws.sort_ws), ws.sort_ws.numel(),
                   ptr(ws.sorted_rows), ptr(ws.sorted_slots), ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp),
                   ptr(self.stamp), ptr(ws.n_owned), s, n_kernels=7)
        self._adam_untouched(ws)

    def embedding_backward(self, ws: _Ws, apply_adam: bool, want_dense_grad: bool, sorted_already: bool = False):
        cfg, E, B, s = self.cfg, self.cfg.emb_dim, ws.B, self._stream()
        K0, K1 = self.K
        KL0, KL1 = self.KLp
        n0, n_l = B * KL0, B * (KL0 + KL1)
        if n_l == 0 or self.total_rows == 0:
            return
        cap = n_l * self.world
        H0 = cfg.hidden[0]
        if not sorted_already:
            self.embedding_sort(ws)
        self.profile_tag = "sharded"
        if self.lazy:
            if want_dense_grad or not apply_adam:
                raise native.TwoTowerNativeError("optimizer_mode='lazy_exact' updates the tables inside the library (attach a FusedAdam)")
            self._call("tt_ebc_backward_adam_lazy", ptr(self.arena), ptr(self.arena_m), ptr(self.arena_v), self.total_rows, E,
                       ptr(ws.sorted_rows), ptr(ws.sorted_slots), cap, ptr(self.row_begin), ptr(self.row_end), ptr(self.row_stamp),
                       ptr(self.stamp), 0, H0 + K0 * E, H0, max(KL0, 1), n0, 0, H0 + K1 * E, H0, max(KL1, 1), n_l - n0, self.act,
                       ptr(ws.piece_ws), ws.piece_ws.numel(), ptr(self.hyper), ptr(ws.n_owned), ptr(ws.dc_table), n_l,
                       ptr(self.colmap[0]), ptr(self.colmap[1]), ptr(self.lazy_ctx), s, n_kernels=3)
            return
        if want_dense_grad and self.arena_grad is None:
            self.arena_grad = torch.zeros_like(self.arena)
        self._call("tt_ebc_backward_adam", ptr(self.arena) if apply_adam else 0, ptr(self.arena_m) if apply_adam else 0,
                   ptr(self.arena_v) if apply_adam else 0, ptr(self.arena_grad) if want_dense_grad else 0,
                   self.total_rows, E, ptr(ws.sorted_rows), ptr(ws.sorted_slots), cap, ptr(self.row_begin), ptr(self.row_end),
                   ptr(self.row_stamp), 0, ptr(self.stamp), 0, H0 + K0 * E, H0, max(KL0, 1), n0, 0, H0 + K1 * E, H0, max(KL1, 1),
                   n_l - n0, self.act, ptr(ws.piece_ws), ws.piece_ws.numel(), ptr(self.hyper),
                   self._table_mode(ws, apply_adam), ptr(ws.n_owned), ptr(ws.dc_table), n_l,
                   ptr(self.colmap_p[0]), ptr(self.colmap_p[1]), s, n_kernels=2)

    def similarity_matrix(self, ws: _Ws) -> torch.Tensor:
        """This rank's rows of the global logits: [B_l, B_g]."""
        B, Bg = ws.B, ws.B * self.world
        S = torch.empty(B, Bg, dtype=torch.float32, device=self.device)
        n, c = ws.e[0].float(), ws.e_all[1].float()
        self._call("tt_similarity_matrix", ptr(n), ptr(c), B, Bg, self.cfg.final_dim, TT_F32, 1.0 / self.cfg.temperature, ptr(S),
                   Bg, self._stream())
        return S

    # ---- full <-> sharded parameter exchange (reference-named state_dict lives on the full tables)
    def _large_index(self, t: int, k: int) -> int:
        return (0 if t == 0 else self.KL[0]) + self.large_keys[t].index(k)

    def load_full_tables(self, full_tables: Sequence[torch.Tensor], which: str = "w"):
        """full_tables: [V_k, E] for notice keys then company keys; small tables are copied whole (replicated), of a large
        table this rank keeps its row shard. `which`: weights "w" or the Adam moments "m" / "v"."""
        self.lazy_mark_current()
        i = 0
        for t in range(2):
            for k in range(self.K[t]):
                tab, view = full_tables[i].to(self.device), self.table_view(t, k, which)
                i += 1
                if self.is_small[t][k]:
                    view.copy_(tab)
                else:
                    view.zero_()
                    lo, hi = self.plan.shard_range(self._large_index(t, k), self.rank)
                    if hi > lo:
                        view[:hi - lo].copy_(tab[lo:hi])

    def gather_full_tables(self, which: str = "w") -> List[torch.Tensor]:
        out = []
        E = self.cfg.emb_dim
        self.flush_lazy()
        if self.total_rows > 0:
            src = {"w": self.arena, "m": self.arena_m, "v": self.arena_v}[which]
            allsh = torch.empty(self.world * src.numel(), dtype=src.dtype, device=self.device)
            self.comm.all_gather(allsh, src.contiguous().view(-1))
            allsh = allsh.view(self.world, self.total_rows, E)
        for t in range(2):
            vocab = self.cfg.tower(t)[2]
            for k in range(self.K[t]):
                if self.is_small[t][k]:
                    out.append(self.table_view(t, k, which).clone())
                else:
                    j = self._large_index(t, k)
                    b, blk = self.plan.local_base[j], self.plan.block[j]
                    out.append(allsh[:, b:b + blk].reshape(-1, E)[:vocab[k]].clone())
        return out


def create_distributed_train_task(notice_keys, company_keys, schema: dict, model: dict, device, comm=None, dtype: str = "bf16",
                                  optimizer_mode: str = "dense_exact", seed_full_init: bool = True, small_tables: bool = True,
                                  bag_len: int = 1):
    """Build the façade on a DistributedEngine. `schema` = {"notice": [[key, vocab], ...], "company": [...]};
    `model` has emb_dim / notice_dense / company_dense / hidden / final_dim / dropout / temperature.
    With seed_full_init every rank draws the full tables with the current torch RNG (same on every rank) and keeps its
    shard, so the initial weights equal a single-GPU model built with the same seed."""
    import torch.nn as nn
    from .modules import TwoTowerModel, TwoTowerTrainTask, _init_like_reference
    if comm is None:
        comm = NcclComm()
    nv = [int(dict(schema["notice"])[k]) for k in notice_keys]
    cv = [int(dict(schema["company"])[k]) for k in company_keys]
    cfg = TwoTowerConfig(list(notice_keys), list(company_keys), nv, cv, model["emb_dim"], model["notice_dense"],
                         model["company_dense"], list(model["hidden"]), model["final_dim"], model.get("dropout", 0.0),
                         model.get("temperature", 1.0), model.get("label_smoothing", 0.0))
    eng = DistributedEngine(cfg, device, comm, optimizer_mode, "sample_major", dtype, small_tables, bag_len)
    # same RNG consumption order as the single-GPU factory: all notice tables, notice dense layers, then company
    full = []
    E = cfg.emb_dim
    for t in range(2):
        name, keys, vocab, dense_dim = cfg.tower(t)
        pre = f"{name}_tower."
        for ki, v in enumerate(vocab):
            if seed_full_init:
                full.append(nn.Embedding(v, E).weight.detach())
            else:
                # tables too large to draw on the host of every rank (10 M rows x 38 tables): N(0, 1) drawn on the device, shard by
                # shard (sharded tables: seeded by table and rank; replicated small tables: by table only, identical everywhere)
                view = eng.table_view(t, ki)
                g = torch.Generator(device=eng.device)
                g.manual_seed(1_000_003 * (t * 4096 + ki) + (0 if eng.is_small[t][ki] else 1 + eng.rank))
                view.normal_(generator=g)
        lin = nn.Linear(dense_dim, cfg.hidden[0])
        eng.dense(pre + "dense_projection.weight").copy_(lin.weight.detach())
        eng.dense(pre + "dense_projection.bias").copy_(lin.bias.detach())
        width, idx = cfg.hidden[0] + len(keys) * E, 0
        for h in cfg.hidden[1:]:
            lin = nn.Linear(width, h)
            eng.dense(pre + f"mlp.{idx}.weight").copy_(lin.weight.detach())
            eng.dense(pre + f"mlp.{idx}.bias").copy_(lin.bias.detach())
            eng.dense(pre + f"mlp.{idx + 2}.weight").fill_(1.0)
            eng.dense(pre + f"mlp.{idx + 2}.bias").zero_()
            width, idx = h, idx + 4
        lin = nn.Linear(width, cfg.final_dim)
        eng.dense(pre + f"mlp.{idx}.weight").copy_(lin.weight.detach())
        eng.dense(pre + f"mlp.{idx}.bias").copy_(lin.bias.detach())
    if seed_full_init:
        eng.load_full_tables(full)
    del full
    model_mod = TwoTowerModel(eng)
    return TwoTowerTrainTask(model_mod, temperature=cfg.temperature, label_smoothing=cfg.label_smoothing)


def load_reference_state_dict(task, sd: Dict[str, torch.Tensor]):
    """Scatter a reference-named (full-table) state_dict into a task built on a DistributedEngine."""
    eng: DistributedEngine = task.engine
    cfg = eng.full_cfg
    full = []
    for t in range(2):
        name, keys, _, _ = cfg.tower(t)
        for k in keys:
            full.append(sd[f"two_tower_model.{name}_tower.categorical_embedder.embeddings.{k}.weight"].to(eng.device))
    eng.load_full_tables(full)
    own = task.state_dict()
    for k, v in sd.items():
        if ".embeddings." in k:
            continue
        own[k].copy_(v.to(eng.device))


def gather_reference_state_dict(task) -> Dict[str, torch.Tensor]:
    """All-gather the shards back into the reference's state_dict names (SURVEY.md §5 checkpoint compatibility)."""
    eng: DistributedEngine = task.engine
    cfg = eng.full_cfg
    tabs = eng.gather_full_tables("w")
    out, i = {}, 0
    for k, v in task.state_dict().items():
        if ".embeddings." in k:
            continue
        out[k] = v.detach().clone()
    for t in range(2):
        name, keys, _, _ = cfg.tower(t)
        for k in keys:
            out[f"two_tower_model.{name}_tower.categorical_embedder.embeddings.{k}.weight"] = tabs[i]
            i += 1
    return out


def _table_param_names(cfg) -> List[str]:
    return [f"two_tower_model.{cfg.tower(t)[0]}_tower.categorical_embedder.embeddings.{k}.weight" for t in range(2) for k in cfg.tower(t)[1]]


def gather_reference_optimizer_state(task) -> dict:
    """torch.optim.Adam-format optimizer state of the whole model: the row-wise shards of the Adam
    moments are all-gathered back into full [V_k, E] tensors, in the parameter order of the reference's
    `train_task.parameters()`, so `torch.optim.Adam.load_state_dict` / a single-GPU FusedAdam can resume from it
    (scripts/train.py:497-526). Collective: every rank calls it."""
    eng: DistributedEngine = task.engine
    opt = task._fused_opt
    sd = opt.state_dict()  # local view: right for the dense parameters and the replicated small tables
    if not sd["state"]:
        return sd
    full = {w: dict(zip(_table_param_names(eng.full_cfg), eng.gather_full_tables(w))) for w in ("m", "v")}
    for i, (name, _p) in enumerate(task.named_parameters()):
        if name in full["m"]:
            sd["state"][i]["exp_avg"] = full["m"][name]
            sd["state"][i]["exp_avg_sq"] = full["v"][name]
    return sd


def load_reference_optimizer_state(task, sd: dict):
    """Inverse of gather_reference_optimizer_state: scatter a full-table torch.optim.Adam state_dict into the shards."""
    eng: DistributedEngine = task.engine
    opt = task._fused_opt
    names = [n for n, _ in task.named_parameters()]
    tabs = {"m": [], "v": []}
    for tn in _table_param_names(eng.full_cfg):
        st = sd["state"].get(names.index(tn)) if sd["state"] else None
        for w, key in (("m", "exp_avg"), ("v", "exp_avg_sq")):
            tabs[w].append(st[key] if st is not None else torch.zeros(1))
    # dense parameters, small tables, counters: the single-engine loader, on a copy whose table entries match the local shapes
    local = {"state": {}, "param_groups": sd["param_groups"], **{k: v for k, v in sd.items() if k not in ("state", "param_groups")}}
    tset = set(_table_param_names(eng.full_cfg))
    for i, (n, p) in enumerate(task.named_parameters()):
        st = sd["state"].get(i) if sd["state"] else None
        if st is None:
            continue
        if n in tset:
            local["state"][i] = {"step": st["step"], "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        else:
            local["state"][i] = st
    opt.load_state_dict(local)
    if sd["state"]:
        eng.load_full_tables(tabs["m"], "m")
        eng.load_full_tables(tabs["v"], "v")
