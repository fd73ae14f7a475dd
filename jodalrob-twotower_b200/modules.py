"""nn.Module façade with the reference's names, signatures and state_dict keys.

Mirrors (paths relative to /root/reference):
  CategoricalEmbedder          src/towers/cat_embed.py:11-190
  BaseTower/Notice/Company     src/towers/tower/{base,notice,company}_tower.py
  TwoTowerModel (+factory)     src/towers/two_tower_model.py:10-165
  TwoTowerTrainTask (+factory) src/towers/two_tower_train_task.py:9-249
so it drops in behind scripts/train.py:207-220,231,285,320-330,368,394,426.  The modules own no
arithmetic: parameters are views into the engine's flat HBM buffers and every forward/backward is a
sequence of C-ABI kernel launches (engine.py).  Two ways to train:
  (i)  autograd-compatible: `loss = task(batch)["loss"]; loss.backward(); torch.optim.Adam.step()` —
       gradients of every parameter (dense [V,E] for tables, like nn.Embedding) come from our kernels;
  (ii) fused: attach `FusedAdam(task, ...)`; then `loss.backward()` runs backward + sort-dedup +
       fused Adam inside the library and `optimizer.step()` / `scheduler.step()` are no-ops, so the
       reference's training loop is unchanged.  `task.fused_step(batch)` is the same without autograd.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from . import native
from .engine import TwoTowerConfig, TwoTowerEngine
from .kjt import KeyedJaggedTensor
from .schema import vocab_sizes_from_metadata

__all__ = ["CategoricalEmbedder", "BaseTower", "NoticeTower", "CompanyTower", "TwoTowerModel", "TwoTowerTrainTask",
           "create_two_tower_model", "create_two_tower_train_task", "FusedAdam", "FusedWarmupScheduler", "GraphedFusedStep"]


class _Table(nn.Module):
    """Stands where the reference has nn.Embedding (cat_embed.py:41-45): `.weight` is a row view of the arena."""

    def __init__(self, weight: torch.Tensor):
        super().__init__()
        self.num_embeddings, self.embedding_dim = weight.shape
        self.weight = nn.Parameter(weight)


class _Linear(nn.Module):
    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.out_features, self.in_features = weight.shape
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)


class _BatchNorm(nn.Module):
    def __init__(self, weight, bias, running_mean, running_var, num_batches_tracked):
        super().__init__()
        self.num_features = weight.shape[0]
        self.eps, self.momentum = 1e-5, 0.1
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)
        self.register_buffer("running_mean", running_mean)
        self.register_buffer("running_var", running_var)
        self.register_buffer("num_batches_tracked", num_batches_tracked)


class _Marker(nn.Module):
    """Parameter-free place holder (ReLU / Dropout slots) so mlp.{0,2,4,...} indices match the reference."""

    def __init__(self, kind: str, p: float = 0.0):
        super().__init__()
        self.kind, self.p = kind, p

    def extra_repr(self):
        return f"{self.kind}" + (f", p={self.p}" if self.kind == "dropout" else "")


def _values_of(kjt) -> torch.Tensor:
    v = kjt.values()
    return v if v.dtype == torch.int64 else v.to(torch.int64)


def _layout_of(kjt) -> Optional[str]:
    return getattr(kjt, "layout", None)


class CategoricalEmbedder(nn.Module):
    def __init__(self, engine: TwoTowerEngine, tower: int):
        super().__init__()
        self._engine = [engine]  # hidden from nn.Module registration
        self._tower = tower
        name, keys, vocab, _ = engine.cfg.tower(tower)
        self.keys = list(keys)
        self.embedding_dim = engine.cfg.emb_dim
        self.vocab_sizes = {k: int(v) for k, v in zip(keys, vocab)}
        self.embeddings = nn.ModuleDict()
        for ki, k in enumerate(self.keys):
            self.embeddings[k] = _Table(engine.table_view(tower, ki))

    def forward(self, kjt, return_dict: bool = True):
        """Pooled lookup of every key (forward only; training goes through TwoTowerTrainTask).
        Raises on malformed input instead of the reference's silent zero fallback (cat_embed.py:180-190)."""
        from . import ops
        eng = self._engine[0]
        if kjt is None:
            raise ValueError("kjt is None")
        values = _values_of(kjt).to(eng.device)
        K = len(self.keys)
        B = values.numel() // K
        out = ops.ebc_forward(eng, self._tower, values, B, layout=_layout_of(kjt) or eng.default_layout)
        if return_dict:
            E = self.embedding_dim
            return {k: out[:, i * E:(i + 1) * E] for i, k in enumerate(self.keys)}
        return out


class BaseTower(nn.Module):
    def __init__(self, engine: TwoTowerEngine, tower: int):
        super().__init__()
        self._engine = [engine]
        self._tower = tower
        cfg = engine.cfg
        name, keys, _, _ = cfg.tower(tower)
        pre = f"{name}_tower."
        self.categorical_keys = list(keys)
        self.device = engine.device
        self.categorical_embedder = CategoricalEmbedder(engine, tower)
        self.dense_projection = _Linear(engine.dense(pre + "dense_projection.weight"), engine.dense(pre + "dense_projection.bias"))
        layers: List[nn.Module] = []
        idx = 0
        for bi, _h in enumerate(cfg.hidden[1:]):
            layers += [_Linear(engine.dense(pre + f"mlp.{idx}.weight"), engine.dense(pre + f"mlp.{idx}.bias")),
                       _Marker("relu"),
                       _BatchNorm(engine.dense(pre + f"mlp.{idx + 2}.weight"), engine.dense(pre + f"mlp.{idx + 2}.bias"),
                                  engine.bn_rm[tower][bi], engine.bn_rv[tower][bi], engine.bn_nbt[tower][bi]),
                       _Marker("dropout", cfg.dropout)]
            idx += 4
        layers.append(_Linear(engine.dense(pre + f"mlp.{idx}.weight"), engine.dense(pre + f"mlp.{idx}.bias")))
        self.mlp = nn.Sequential(*layers)

    def forward(self, tower_input: Dict[str, torch.Tensor]) -> torch.Tensor:
        """[B, final_embedding_dim] L2-normalised embeddings (base_tower.py:101-147), forward only."""
        eng = self._engine[0]
        dense = tower_input["dense"].to(eng.device)
        kjt = tower_input["kjt"].to(eng.device)
        values = _values_of(kjt)
        B = dense.shape[0]
        ws = eng.workspace(B)
        lay = native.TT_LAYOUT_KEY_MAJOR if (_layout_of(kjt) or eng.default_layout) == "key_major" else native.TT_LAYOUT_SAMPLE_MAJOR
        eng._check_inputs(self._tower, dense, values, B)
        with torch.no_grad():
            return eng.tower_forward(self._tower, ws, dense, values, self.training, lay).clone()


class NoticeTower(BaseTower):
    def __init__(self, engine: TwoTowerEngine):
        super().__init__(engine, 0)


class CompanyTower(BaseTower):
    def __init__(self, engine: TwoTowerEngine):
        super().__init__(engine, 1)


class TwoTowerModel(nn.Module):
    def __init__(self, engine: TwoTowerEngine):
        super().__init__()
        self._engine = [engine]
        self.device = engine.device
        self.final_embedding_dim = engine.cfg.final_dim
        self.notice_tower = NoticeTower(engine)
        self.company_tower = CompanyTower(engine)

    @property
    def engine(self) -> TwoTowerEngine:
        return self._engine[0]

    def forward(self, notice_input, company_input, return_similarity: bool = False, temperature: float = 1.0):
        n = self.notice_tower(notice_input)
        c = self.company_tower(company_input)
        if n.size(0) != c.size(0):
            raise ValueError(f"Notice와 Company 배치 크기가 다릅니다: {n.size(0)} vs {c.size(0)}")
        if return_similarity:
            from . import ops
            return {"notice_embeddings": n, "company_embeddings": c,
                    "similarity_matrix": ops.similarity_matrix(self.engine, n, c, 1.0 / temperature)}
        return n, c

    def get_notice_embeddings(self, notice_input):
        return self.notice_tower(notice_input)

    def get_company_embeddings(self, company_input):
        return self.company_tower(company_input)

    def compute_similarity(self, notice_emb, company_emb, temperature: float = 1.0):
        from . import ops
        return ops.similarity_matrix(self.engine, notice_emb, company_emb, 1.0 / temperature)


class _LazyResult(dict):
    """Result dict of TwoTowerTrainTask.forward; the B x B `similarity_matrix` (two_tower_train_task.py:94) is
    only materialised if somebody reads it — the train step itself never stores the logits."""

    def __init__(self, base: dict, make_sim):
        super().__init__(base)
        self._make_sim = make_sim

    def __getitem__(self, k):
        if k == "similarity_matrix" and not dict.__contains__(self, k):
            dict.__setitem__(self, k, self._make_sim())
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        return self[k] if (k == "similarity_matrix" or dict.__contains__(self, k)) else default

    def __contains__(self, k):
        return k == "similarity_matrix" or dict.__contains__(self, k)

    def keys(self):
        return list(dict.keys(self)) + ([] if dict.__contains__(self, "similarity_matrix") else ["similarity_matrix"])


class _TrainStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, task, nd, nv, cd, cv, layout, masks, *params):
        eng = task.engine
        if task._fused_opt is not None:
            o = task._fused_opt
            eng.hyper_advance(o.lr, o.warmup_steps, o.betas, o.eps, o.weight_decay)
        elif eng.lazy:
            raise native.TwoTowerNativeError("optimizer_mode='lazy_exact' needs a FusedAdam(task, ...): the table update runs inside the library")
        elif eng.cfg.dropout > 0:
            eng.hyper_advance(0.0, 0, (0.9, 0.999), 1e-8, 0.0)  # only the dropout RNG step matters here
        eng._lazy_offset = -1 if task._fused_opt is not None else 0
        try:
            out5, ws = eng.forward(nd, nv, cd, cv, True, layout, masks)
        finally:
            eng._lazy_offset = 0
        ctx.task, ctx.ws = task, ws
        res = out5.clone()
        ctx.mark_non_differentiable(res)
        return res[0].clone(), res

    @staticmethod
    def backward(ctx, dloss, _dres):
        task, ws = ctx.task, ctx.ws
        eng = task.engine
        if task._fused_opt is not None:
            # backward + sort-dedup + fused Adam all inside the library. The upstream gradient scales every gradient, like under
            # autograd ((loss / accum).backward(), a GradScaler, loss + other terms). Contract of this mode: exactly one
            # backward per train-mode forward — the optimizer scalars (Adam step, lr schedule, dropout counter) advance in forward.
            eng.backward(ws, float(dloss))
            eng.reduce_dense_grads()
            eng.embedding_backward(ws, apply_adam=True, want_dense_grad=False)
            eng.dense_adam()
            return (None,) * 8  # 7 plain arguments + the anchor
        eng.backward(ws, float(dloss))
        eng.embedding_backward(ws, apply_adam=False, want_dense_grad=True)
        grads = []
        for kind, ref in task._param_src:
            if kind == "table":
                grads.append(eng.table_view(ref[0], ref[1], "g").clone())
            else:
                grads.append(eng.dense(ref, "g").clone())
        return (None,) * 7 + tuple(grads)


def _cosine_embedding_loss(s: torch.Tensor) -> torch.Tensor:
    """two_tower_train_task.py:136-158 verbatim in meaning: diagonal sims with target +1, off-diagonal with target -1, through the
    (degenerate, 1-D) F.cosine_embedding_loss against a vector of ones."""
    import torch.nn.functional as F
    b = s.size(0)
    off = ~torch.eye(b, dtype=torch.bool, device=s.device)
    sims = torch.cat([torch.diag(s), s[off]])
    tg = torch.cat([torch.ones(b, device=s.device), -torch.ones(b * b - b, device=s.device)])
    return F.cosine_embedding_loss(sims.unsqueeze(1), torch.ones_like(sims.unsqueeze(1)), tg)


class _CosineStepFn(torch.autograd.Function):
    """Train step with loss_type="cosine_embedding" (not the default, off the fast path): the towers, their
    backward and the optimizer run in the library; the loss on the materialised [B,B] matrix and dS -> (dN, dC) are torch CUDA
    ops. Its gradient is numerically ~0 by construction (d sign(s)/ds), exactly like the reference's."""

    @staticmethod
    def forward(ctx, task, nd, nv, cd, cv, layout, masks, *params):
        eng = task.engine
        if task._fused_opt is not None:
            o = task._fused_opt
            eng.hyper_advance(o.lr, o.warmup_steps, o.betas, o.eps, o.weight_decay)
        elif eng.lazy:
            raise native.TwoTowerNativeError("optimizer_mode='lazy_exact' needs a FusedAdam(task, ...): the table update runs inside the library")
        elif eng.cfg.dropout > 0:
            eng.hyper_advance(0.0, 0, (0.9, 0.999), 1e-8, 0.0)
        eng._lazy_offset = -1 if task._fused_opt is not None else 0
        try:
            out5, ws = eng.forward(nd, nv, cd, cv, True, layout, masks)
        finally:
            eng._lazy_offset = 0
        s = eng.similarity_matrix(ws)
        res = out5.clone()
        res[0] = _cosine_embedding_loss(s)
        ctx.task, ctx.ws, ctx.s = task, ws, s
        ctx.mark_non_differentiable(res)
        return res[0].clone(), res

    @staticmethod
    def backward(ctx, dloss, _dres):
        task, ws = ctx.task, ctx.ws
        eng = task.engine
        with torch.enable_grad():
            s_ = ctx.s.detach().requires_grad_(True)
            (ds,) = torch.autograd.grad(_cosine_embedding_loss(s_), s_)
        fused = task._fused_opt is not None
        scale = float(dloss) / eng.cfg.temperature
        n, c = ws.e[0].float(), ws.e[1].float()
        ws.de[0].copy_(torch.mm(ds, c) * scale)
        ws.de[1].copy_(torch.mm(ds.t(), n) * scale)
        eng.backward(ws, 1.0, de_given=True)
        if fused:
            eng.reduce_dense_grads()
            eng.embedding_backward(ws, apply_adam=True, want_dense_grad=False)
            eng.dense_adam()
            return (None,) * 8
        eng.embedding_backward(ws, apply_adam=False, want_dense_grad=True)
        grads = []
        for kind, ref in task._param_src:
            grads.append(eng.table_view(ref[0], ref[1], "g").clone() if kind == "table" else eng.dense(ref, "g").clone())
        return (None,) * 7 + tuple(grads)


class TwoTowerTrainTask(nn.Module):
    def __init__(self, two_tower_model: TwoTowerModel, temperature: float = 1.0, loss_type: str = "cross_entropy",
                 label_smoothing: float = 0.0):
        super().__init__()
        if loss_type not in ["cross_entropy", "cosine_embedding"]:
            raise ValueError(f"Unsupported loss_type: {loss_type}")
        self.two_tower_model = two_tower_model
        self.temperature = temperature
        self.loss_type = loss_type
        self.label_smoothing = label_smoothing
        self._fused_opt: Optional["FusedAdam"] = None
        self._anchor = torch.zeros((), device=two_tower_model.device, requires_grad=True)
        self.debug_pair_check = False
        self._index_params()

    @property
    def engine(self) -> TwoTowerEngine:
        return self.two_tower_model.engine

    def _index_params(self):
        eng = self.engine
        self._param_list, self._param_src = [], []
        for name, p in self.named_parameters():
            self._param_list.append(p)
            short = name[len("two_tower_model."):]
            if ".embeddings." in short:
                tower = 0 if short.startswith("notice") else 1
                key = short.split(".embeddings.")[1].rsplit(".weight", 1)[0]
                keys = eng.cfg.notice_keys if tower == 0 else eng.cfg.company_keys
                self._param_src.append(("table", (tower, keys.index(key))))
            else:
                self._param_src.append(("dense", short))

    def _apply(self, fn, *a, **kw):
        # parameters are views of the engine's flat buffers; moving/casting them would break the aliasing
        probe = fn(torch.zeros(1, device=self.engine.device))
        if probe.device != self.engine.device or probe.dtype != torch.float32:
            raise RuntimeError("TwoTowerTrainTask lives on the device it was created on (fp32 masters); "
                               "pass device= to create_two_tower_train_task instead of calling .to()/.half()")
        return self

    def state_dict(self, *args, **kwargs):
        self.engine.flush_lazy()  # lazy exact-replay Adam: every row up to date before anybody reads the tables
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        out = super().load_state_dict(state_dict, *args, **kwargs)
        self.engine.lazy_mark_current()
        return out

    def _sync_cfg(self):
        cfg = self.engine.cfg
        cfg.temperature, cfg.label_smoothing = float(self.temperature), float(self.label_smoothing)

    def _unpack(self, batch):
        eng = self.engine
        n, c = batch["notice"], batch["company"]
        nb, cb = n["dense"].size(0), c["dense"].size(0)
        if nb != cb:
            raise ValueError(f"Notice와 Company 배치 크기가 다릅니다: {nb} vs {cb}")
        nd, cd = n["dense"].to(eng.device, non_blocking=True), c["dense"].to(eng.device, non_blocking=True)
        nk, ck = n["kjt"].to(eng.device), c["kjt"].to(eng.device)
        lay = _layout_of(nk) or eng.default_layout
        return nd, _values_of(nk), cd, _values_of(ck), lay

    def forward(self, batch, return_metrics: bool = False, dropout_masks=None):
        self._sync_cfg()
        eng = self.engine
        nd, nv, cd, cv, lay = self._unpack(batch)
        self._last_B = nd.shape[0]
        if self.loss_type != "cross_entropy":
            return self._forward_cosine_embedding(nd, nv, cd, cv, lay, return_metrics, dropout_masks)
        if self.training and torch.is_grad_enabled():
            params = (self._anchor,) if self._fused_opt is not None else tuple(self._param_list)
            loss, res = _TrainStepFn.apply(self, nd, nv, cd, cv, lay, dropout_masks, *params)
            ws = eng.workspace(nd.shape[0])
        else:
            with torch.no_grad():
                out5, ws = eng.forward(nd, nv, cd, cv, self.training, lay, dropout_masks)
                res = out5.clone()
                loss = res[0]
        if self.debug_pair_check and self.training and not hasattr(self, "_pair_check_done"):
            # the reference runs this once, at its first forward (two_tower_train_task.py:80-83); here it is opt-in
            # (task.debug_pair_check = True) because it materialises the B x B matrix the train path otherwise never stores
            self._verify_positive_pair_alignment(eng.similarity_matrix(ws))
            self._pair_check_done = True
        if not return_metrics:
            return loss
        det = res.detach()
        return _LazyResult({"loss": loss, "accuracy": det[1], "positive_similarity_mean": det[2],
                            "negative_similarity_mean": det[3], "similarity_gap": det[4]},
                           lambda: eng.similarity_matrix(ws))

    def fused_step(self, batch, lr=None, warmup_steps=None) -> torch.Tensor:
        """One whole train step inside the library, no autograd. Returns the device tensor
        [loss, accuracy, positive_similarity_mean, negative_similarity_mean, similarity_gap]."""
        self._sync_cfg()
        o = self._fused_opt
        if o is None:
            raise RuntimeError("attach a FusedAdam(task, ...) before fused_step")
        if self.loss_type != "cross_entropy":  # non-default loss: library towers/optimizer around a torch-evaluated loss
            if lr is not None or warmup_steps is not None:
                raise ValueError("per-call lr/warmup overrides are only supported for loss_type='cross_entropy'")
            was = self.training
            self.train()
            out = self(batch, return_metrics=True)
            out["loss"].backward()
            self.train(was)
            return torch.stack([out[k].detach().float() for k in ("loss", "accuracy", "positive_similarity_mean",
                                                                    "negative_similarity_mean", "similarity_gap")])
        nd, nv, cd, cv, lay = self._unpack(batch)
        return self.engine.fused_step(nd, nv, cd, cv, lr if lr is not None else o.lr,
                                      warmup_steps if warmup_steps is not None else o.warmup_steps, o.betas, o.eps,
                                      o.weight_decay, lay)

    def _forward_cosine_embedding(self, nd, nv, cd, cv, lay, return_metrics, dropout_masks=None):
        # two_tower_train_task.py:136-158 — the non-default, near-zero-gradient alternative loss, evaluated from the
        # materialised matrix. Training goes through _CosineStepFn; not on the hot path.
        eng = self.engine
        if self.training and torch.is_grad_enabled():
            params = (self._anchor,) if self._fused_opt is not None else tuple(self._param_list)
            loss, res = _CosineStepFn.apply(self, nd, nv, cd, cv, lay, dropout_masks, *params)
            ws = eng.workspace(nd.shape[0])
            s = None
        else:
            with torch.no_grad():
                res, ws = eng.forward(nd, nv, cd, cv, self.training, lay, dropout_masks)
                res = res.clone()
                s = eng.similarity_matrix(ws)
                loss = _cosine_embedding_loss(s)
        if not return_metrics:
            return loss
        r = res.detach()
        return _LazyResult({"loss": loss, "accuracy": r[1], "positive_similarity_mean": r[2], "negative_similarity_mean": r[3],
                            "similarity_gap": r[4]}, (lambda: s) if s is not None else (lambda: eng.similarity_matrix(ws)))

    def eval_ranks(self, batch=None):
        """(rank i32 [B], positive logit f32 [B]): position of every notice's own company in a stable descending sort of its
        row of S, from the tower embeddings of the last forward (or of `batch`, forwarded in the current mode) — the B x B
        matrix is never stored. recall@k / MRR follow from it (evaluator.py:20-71)."""
        from .evaluator import ranks_from_embeddings
        if batch is not None:
            with torch.no_grad():
                self(batch)
        if getattr(self, "_last_B", None) is None:
            raise RuntimeError("eval_ranks() needs a forward first (or pass a batch)")
        ws = self.engine.workspace(self._last_B)
        return ranks_from_embeddings(ws.e[0], ws.e[1], 1.0 / self.temperature)

    def predict_batch(self, batch, top_k: int = 10):
        self.eval()
        with torch.no_grad():
            n, c = self.two_tower_model(batch["notice"], batch["company"])
            from . import ops
            s = ops.similarity_matrix(self.engine, n, c, 1.0 / self.temperature)
            top_sim, top_idx = torch.topk(s, k=top_k, dim=1)
            return {"top_similarities": top_sim, "top_indices": top_idx, "all_similarities": s}

    def _verify_positive_pair_alignment(self, similarity_matrix: torch.Tensor):
        """One-time diagnostic of the first train-mode forward (two_tower_train_task.py:80-83, 253-290): is the diagonal really the
        positive? Top-1 hit rate of the diagonal along rows (notice -> company) and columns (company -> notice), the share of rows
        whose diagonal IS the row maximum, and the reference's three-way verdict (both hit rates < 0.05: the loader's notice /
        company order is broken; either < 0.3: too few positives; else fine). The numbers are kept in `self.pair_alignment`."""
        b = similarity_matrix.size(0)
        idx = torch.arange(b, device=similarity_matrix.device)
        row_hit = (similarity_matrix.argmax(dim=1) == idx).float().mean().item()
        col_hit = (similarity_matrix.argmax(dim=0) == idx).float().mean().item()
        diag_is_max = (similarity_matrix.diag() == similarity_matrix.max(dim=1)[0]).float().mean().item()
        if row_hit < 0.05 and col_hit < 0.05:
            verdict = "CRITICAL"
        elif row_hit < 0.3 or col_hit < 0.3:
            verdict = "WARNING"
        else:
            verdict = "OK"
        self.pair_alignment = {"row_top1": row_hit, "col_top1": col_hit, "diag_is_row_max": diag_is_max, "verdict": verdict}
        print("[Positive Pair Alignment Check]")
        print(f"   Notice->Company top-1 accuracy: {row_hit:.3f}")
        print(f"   Company->Notice top-1 accuracy: {col_hit:.3f}")
        print(f"   share of rows whose diagonal is the row maximum: {diag_is_max:.3f}")
        if verdict == "CRITICAL":
            print("   CRITICAL: positive pairs are not aligned - notice / company order differs in the DataLoader;")
            print("   build the pairs with one sampler / one index")
        elif verdict == "WARNING":
            print("   WARNING: weak positive-pair alignment - check the share of positives per batch")
        else:
            print("   positive-pair alignment looks fine")
        print()


class GraphedFusedStep:
    """The whole fused train step captured once as a CUDA graph (the step is launch-bound at the reference's model size:
    ~85 kernels of a few microseconds each). Every per-step scalar (Adam step, lr schedule, bias corrections, dropout
    counter, dedup stamp) lives on the device, so replaying the graph IS a new optimizer step.

        step = GraphedFusedStep(task, example_batch)      # needs FusedAdam(task, ...) attached
        for batch in loader: out5 = step(batch)           # [loss, acc, pos_mean, neg_mean, gap] on the device
    """

    def __init__(self, task: "TwoTowerTrainTask", example_batch, compact_inputs: bool = False):
        """compact_inputs: with a bf16 engine the dense features are staged in bf16 — the engine's first act on them is the
        round-to-nearest-even cast to bf16 anyway, so the results are bit-identical, the host->device copy halves (12.6 -> 6.3 MB at
        B = 8192) and the two cast kernels drop out of the step. `pack()` / `load()` convert; the ids stay int64."""
        if task._fused_opt is None:
            raise RuntimeError("attach a FusedAdam(task, ...) before capturing the fused step")
        if task.loss_type != "cross_entropy":
            raise ValueError("GraphedFusedStep captures the cross_entropy hot path; loss_type='cosine_embedding' runs eagerly")
        self.task = task
        eng = task.engine
        task._sync_cfg()
        nd, nv, cd, cv, lay = task._unpack(example_batch)
        self.layout = lay
        self.compact = bool(compact_inputs and eng.dtype == "bf16")
        if self.compact:
            nd, cd = nd.to(torch.bfloat16), cd.to(torch.bfloat16)
        # the four inputs live back to back in ONE buffer, so a pre-packed batch is staged with a single copy
        self._slots, off = [], 0
        for t in (nd, nv, cd, cv):
            nbytes = t.numel() * t.element_size()
            self._slots.append((off, nbytes, t.dtype, tuple(t.shape)))
            off += (nbytes + 255) // 256 * 256
        self.static_blob = torch.empty(max(off, 256), dtype=torch.uint8, device=eng.device)
        self.static = [self.static_blob[o:o + n].view(dt).view(shape) for o, n, dt, shape in self._slots]
        for s, t in zip(self.static, (nd, nv, cd, cv)):
            s.copy_(t)
        o = task._fused_opt
        self._args = (o.lr, o.warmup_steps, o.betas, o.eps, o.weight_decay)
        hyper0 = eng.hyper.clone()
        lazy0 = (eng.row_last.clone(), eng.lazy_hist.clone(), eng._steps_since_flush) if eng.lazy else None
        state0 = (eng.flat_w.clone(), eng.flat_m.clone(), eng.flat_v.clone(),
                  [[b.clone() for b in bl] for bl in eng.bn_rm], [[b.clone() for b in bl] for bl in eng.bn_rv],
                  [[b.clone() for b in bl] for bl in eng.bn_nbt])
        side = torch.cuda.Stream(eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side):  # warm-up (allocates the workspace, sets kernel attributes)
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(eng.device).wait_stream(side)
        torch.cuda.synchronize(eng.device)
        self.graph = torch.cuda.CUDAGraph()
        l0 = eng.launches
        with torch.cuda.graph(self.graph):
            self.out = self._run()
        self.launches_per_step = eng.launches - l0
        # capture and warm-up must not count as training: restore the parameters and optimizer state. The dedup epoch
        # (eng.stamp) is NOT rolled back: the warm-up left row_stamp / row_begin / row_end entries carrying its values, and an
        # epoch that repeats would make a later step mistake those rows for rows of its own batch. It only has to be monotonic.
        torch.cuda.synchronize(eng.device)
        eng.hyper.copy_(hyper0)
        if lazy0 is not None:
            eng.row_last.copy_(lazy0[0]); eng.lazy_hist.copy_(lazy0[1]); eng._steps_since_flush = lazy0[2]
        eng.flat_w.copy_(state0[0]); eng.flat_m.copy_(state0[1]); eng.flat_v.copy_(state0[2])
        for dst, src in ((eng.bn_rm, state0[3]), (eng.bn_rv, state0[4]), (eng.bn_nbt, state0[5])):
            for dl, sl in zip(dst, src):
                for d, s_ in zip(dl, sl):
                    d.copy_(s_)

    def _run(self):
        nd, nv, cd, cv = self.static
        return self.task.engine.fused_step(nd, nv, cd, cv, *self._args, self.layout)

    def load(self, batch):
        """Stage a batch into the graph's static input buffers (device-to-device or host-to-device, async)."""
        n, c = batch["notice"], batch["company"]
        self.static[0].copy_(n["dense"], non_blocking=True)
        self.static[1].copy_(n["kjt"].values(), non_blocking=True)
        self.static[2].copy_(c["dense"], non_blocking=True)
        self.static[3].copy_(c["kjt"].values(), non_blocking=True)

    def pack(self, batch, pin: bool = False) -> torch.Tensor:
        """The batch as one uint8 tensor in the layout of the static input buffer (same device as the batch; pinned host
        memory with pin=True). `step(packed=...)` stages it with a single asynchronous copy."""
        n, c = batch["notice"], batch["company"]
        parts = (n["dense"], n["kjt"].values(), c["dense"], c["kjt"].values())
        out = torch.empty(self.static_blob.numel(), dtype=torch.uint8, device=parts[0].device, pin_memory=pin and parts[0].device.type == "cpu")
        for (o, nb, dt, shape), t in zip(self._slots, parts):
            out[o:o + nb].view(dt).view(shape).copy_(t.to(dt).reshape(shape))
        return out

    def stage_packed(self, packed: torch.Tensor):
        """Stage a pre-packed batch (see pack()) into the static input buffer without replaying: one asynchronous copy on the current
        stream. The source buffer may be refilled as soon as this copy has run (record an event behind it); `step()` then replays."""
        self.static_blob.copy_(packed, non_blocking=True)

    def __call__(self, batch=None, packed: Optional[torch.Tensor] = None) -> torch.Tensor:
        if packed is not None:
            self.static_blob.copy_(packed, non_blocking=True)
        elif batch is not None:
            self.load(batch)
        self.task.engine.note_step()  # lazy exact-replay Adam: flush (eagerly, between replays) before the history ring wraps
        self.graph.replay()
        self.task.engine.launches += self.launches_per_step
        return self.out


class FusedAdam:
    """Optimizer shim for the fused path: the Adam update (coupled L2 weight decay, bias correction, the
    LambdaLR warm-up of scripts/train.py:235-242) runs inside the library during loss.backward();
    zero_grad()/step() keep the reference loop (scripts/train.py:319-330) unchanged."""

    def __init__(self, task: TwoTowerTrainTask, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, warmup_steps: int = 0):
        self.task, self.lr, self.betas, self.eps, self.weight_decay = task, lr, tuple(betas), eps, weight_decay
        self.warmup_steps = int(warmup_steps)
        self.param_groups = [{"lr": lr, "betas": self.betas, "eps": eps, "weight_decay": weight_decay,
                              "params": task._param_list}]
        task._fused_opt = self

    def zero_grad(self, set_to_none: bool = True):
        return None

    def step(self, closure=None):
        return None

    # ---- checkpointing in torch.optim.Adam's own format: the reference saves `optimizer.state_dict()` next to
    # `model.state_dict()` and loads both back to resume (scripts/train.py:497-512, 523-526). A checkpoint written through the
    # fused path can be resumed by the reference's torch.optim.Adam (and by façade mode (i)), and the other way round.
    def _moment_views(self):
        eng = self.task.engine
        for kind, ref in self.task._param_src:
            if kind == "table":
                yield eng.table_view(ref[0], ref[1], "m"), eng.table_view(ref[0], ref[1], "v")
            else:
                yield eng.dense(ref, "m"), eng.dense(ref, "v")

    def _next_lr(self, sched_step: int) -> float:
        w = self.warmup_steps
        return self.lr * (sched_step / w if (w > 0 and sched_step < w) else 1.0)

    def state_dict(self):
        self.task.engine.flush_lazy()
        h = self.task.engine.hyper_host()
        step = int(h["step"])
        state = {}
        if step > 0:
            for i, (m, v) in enumerate(self._moment_views()):
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().clone(), "exp_avg_sq": v.detach().clone()}
        group = {"lr": self._next_lr(int(h["sched_step"])), "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "initial_lr": self.lr, "params": list(range(len(self.task._param_list)))}
        # extra top-level keys are ignored by torch.optim.Optimizer.load_state_dict; they make our own resume exact
        return {"state": state, "param_groups": [group],
                "tt_fused": {"sched_step": int(h["sched_step"]), "warmup_steps": self.warmup_steps, "rng_step": int(h["rng_step"])}}

    def load_state_dict(self, sd):
        eng = self.task.engine
        if "param_groups" not in sd:  # the pre-checkpoint-compat format of this class
            for k in ("hyper", "flat_m", "flat_v"):
                getattr(eng, k).copy_(sd[k])
            c = sd["config"]
            self.lr, self.betas, self.eps, self.weight_decay = c["lr"], tuple(c["betas"]), c["eps"], c["weight_decay"]
            self.warmup_steps = c["warmup_steps"]
            return
        g = sd["param_groups"][0]
        if len(g["params"]) != len(self.task._param_list):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
        self.lr = float(g.get("initial_lr", g["lr"]))
        self.betas, self.eps, self.weight_decay = tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"])
        state, step = sd["state"], 0
        for i, (m, v) in enumerate(self._moment_views()):
            st = state.get(i, state.get(str(i)))
            if st is None:
                m.zero_()
                v.zero_()
                continue
            m.copy_(st["exp_avg"].to(m.device))
            v.copy_(st["exp_avg_sq"].to(v.device))
            step = max(step, int(float(st["step"])))
        extra = sd.get("tt_fused", {})
        if "warmup_steps" in extra:
            self.warmup_steps = int(extra["warmup_steps"])
        # a reference checkpoint carries no scheduler counter here (LambdaLR has its own state_dict): scheduler.step() was
        # called once per optimizer step in its loop (scripts/train.py:327-330), so the two counters are equal
        sched = int(extra.get("sched_step", step))
        h64 = eng.hyper.view(torch.int64)
        h64[0] = step
        h64[1] = sched
        if "rng_step" in extra:
            h64[7] = int(extra["rng_step"])
        self.param_groups[0].update(lr=self._next_lr(sched), betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)
        eng.lazy_mark_current()


class FusedWarmupScheduler:
    """Stand-in for LambdaLR(optimizer, lr_lambda) of scripts/train.py:242; the schedule lives on the device."""

    def __init__(self, optimizer: FusedAdam, warmup_steps: int):
        self.optimizer = optimizer
        optimizer.warmup_steps = int(warmup_steps)

    def step(self):
        return None

    def get_last_lr(self):
        return [self.optimizer.task.engine.hyper_host()["lr"]]

    def state_dict(self):
        """LambdaLR-shaped (loadable by torch.optim.lr_scheduler.LambdaLR.load_state_dict)."""
        n = int(self.optimizer.task.engine.hyper_host()["sched_step"])
        lr = self.optimizer._next_lr(n)
        return {"base_lrs": [self.optimizer.lr], "last_epoch": n, "_step_count": n + 1, "_get_lr_called_within_step": False,
                "_last_lr": [lr], "lr_lambdas": [None], "_is_initial": False}

    def load_state_dict(self, sd):
        self.optimizer.task.engine.hyper.view(torch.int64)[1] = int(sd["last_epoch"])


def _init_like_reference(engine: TwoTowerEngine):
    """Same initial distribution AND the same RNG consumption order as the reference constructors
    (nn.Embedding N(0,1) per key, then nn.Linear defaults, BatchNorm gamma=1/beta=0; notice tower first),
    so torch.manual_seed(s) before the factory reproduces a seeded reference bit-for-bit."""
    cfg = engine.cfg
    for t in range(2):
        name, keys, vocab, dense_dim = cfg.tower(t)
        pre = f"{name}_tower."
        for ki, v in enumerate(vocab):
            engine.table_view(t, ki).copy_(nn.Embedding(v, cfg.emb_dim).weight.detach())
        lin = nn.Linear(dense_dim, cfg.hidden[0])
        engine.dense(pre + "dense_projection.weight").copy_(lin.weight.detach())
        engine.dense(pre + "dense_projection.bias").copy_(lin.bias.detach())
        width, idx = cfg.hidden[0] + len(keys) * cfg.emb_dim, 0
        for h in cfg.hidden[1:]:
            lin = nn.Linear(width, h)
            engine.dense(pre + f"mlp.{idx}.weight").copy_(lin.weight.detach())
            engine.dense(pre + f"mlp.{idx}.bias").copy_(lin.bias.detach())
            engine.dense(pre + f"mlp.{idx + 2}.weight").fill_(1.0)
            engine.dense(pre + f"mlp.{idx + 2}.bias").zero_()
            width, idx = h, idx + 4
        lin = nn.Linear(width, cfg.final_dim)
        engine.dense(pre + f"mlp.{idx}.weight").copy_(lin.weight.detach())
        engine.dense(pre + f"mlp.{idx}.bias").copy_(lin.bias.detach())


def create_two_tower_model(notice_categorical_keys: Sequence[str], company_categorical_keys: Sequence[str],
                           metadata_path: str = "meta/metadata.csv", categorical_embedding_dim: int = 64,
                           notice_dense_input_dim: int = 256, company_dense_input_dim: int = 128,
                           tower_hidden_dims: Optional[List[int]] = None, final_embedding_dim: int = 128,
                           dropout_rate: float = 0.2, device="cuda:0", vocab_sizes: Optional[Dict[str, Dict[str, int]]] = None,
                           optimizer_mode: str = "dense_exact", value_layout: str = "sample_major",
                           dtype: str = "fp32") -> TwoTowerModel:
    if tower_hidden_dims is None:
        tower_hidden_dims = [256, 128]
    nk, ck = list(notice_categorical_keys), list(company_categorical_keys)
    if vocab_sizes is None:
        vocab_sizes = {"notice": vocab_sizes_from_metadata(metadata_path, "notice", nk),
                       "company": vocab_sizes_from_metadata(metadata_path, "company", ck)}
    cfg = TwoTowerConfig(nk, ck, [int(vocab_sizes["notice"][k]) for k in nk], [int(vocab_sizes["company"][k]) for k in ck],
                         categorical_embedding_dim, notice_dense_input_dim, company_dense_input_dim,
                         list(tower_hidden_dims), final_embedding_dim, dropout_rate)
    engine = TwoTowerEngine(cfg, device, optimizer_mode, value_layout, dtype)
    _init_like_reference(engine)
    return TwoTowerModel(engine)


def create_two_tower_train_task(notice_categorical_keys, company_categorical_keys, metadata_path: str = "meta/metadata.csv",
                                categorical_embedding_dim: int = 64, notice_dense_input_dim: int = 256,
                                company_dense_input_dim: int = 128, tower_hidden_dims=None, final_embedding_dim: int = 128,
                                dropout_rate: float = 0.2, temperature: float = 1.0, loss_type: str = "cross_entropy",
                                device="cuda:0", label_smoothing: float = 0.0, vocab_sizes=None,
                                optimizer_mode: str = "dense_exact", value_layout: str = "sample_major",
                                dtype: str = "fp32") -> TwoTowerTrainTask:
    model = create_two_tower_model(notice_categorical_keys, company_categorical_keys, metadata_path, categorical_embedding_dim,
                                   notice_dense_input_dim, company_dense_input_dim, tower_hidden_dims, final_embedding_dim,
                                   dropout_rate, device, vocab_sizes, optimizer_mode, value_layout, dtype)
    return TwoTowerTrainTask(model, temperature=temperature, loss_type=loss_type, label_smoothing=label_smoothing)
