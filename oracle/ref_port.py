"""TEST INFRASTRUCTURE ONLY — torch-CPU port of the reference two-tower train step.

Not part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file,
and only as the checker / the timed CPU baseline.  The product path
(``jodalrob_twotower_b200``) never imports anything under ``oracle/``.

Why a port: the reference is pure Python/PyTorch and cannot travel to the GPU box
(``/root/reference`` does not exist there).  This file restates it with the same
modules, the same op order and the same ``state_dict`` names, and is pinned
against the unmodified reference by ``oracle/make_golden.py`` →
``tests/golden/*.npz`` and ``tests/test_oracle_vs_golden.py`` (bit-exact on
CPU in the build container).

Reference anchors (paths relative to /root/reference):
  * vocab sizes                     src/towers/cat_embed.py:50-85   (count + 10, default 1000)
  * KJT decode + clamp              src/towers/cat_embed.py:88-123  (sample-major, lengths ignored)
  * per-key nn.Embedding + concat   src/towers/cat_embed.py:41-45,158-177
  * tower                           src/towers/tower/base_tower.py:59,71-99,133-145
  * model                           src/towers/two_tower_model.py:43-100
  * logits / loss / metrics         src/towers/two_tower_train_task.py:99-179
  * optimizer + warm-up             scripts/train.py:231-242,319-330
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class PortKJT:
    """values()/lengths()/keys() container in the reference's sample-major layout
    (src/towers/pairs/unified_bid_data_loader.py:827-841)."""

    def __init__(self, keys: Sequence[str], values: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        self._k = list(keys)
        self._v = values
        self._l = lengths if lengths is not None else torch.ones_like(values)

    def keys(self):
        return self._k

    def values(self):
        return self._v

    def lengths(self):
        return self._l

    def to(self, device, non_blocking=False):
        return PortKJT(self._k, self._v.to(device), self._l.to(device))


class PortEmbedder(nn.Module):
    def __init__(self, keys: Sequence[str], vocab: Dict[str, int], dim: int, decode: str = "vectorized"):
        super().__init__()
        self.keys = list(keys)
        self.vocab_sizes = {k: int(vocab[k]) for k in self.keys}
        self.embedding_dim = dim
        self.decode = decode
        self.embeddings = nn.ModuleDict({k: nn.Embedding(self.vocab_sizes[k], dim) for k in self.keys})

    def decode_ids(self, values: torch.Tensor) -> List[torch.Tensor]:
        nkeys = len(self.keys)
        bsz = len(values) // nkeys
        if self.decode == "loop":
            # the reference's own scalar loop, cat_embed.py:103-121 — this is where 70 % of
            # its CPU step goes; kept for the "as-is" CPU baseline
            out = []
            for ki, key in enumerate(self.keys):
                vmax = self.vocab_sizes[key] - 1
                col = [torch.clamp(values[b * nkeys + ki], min=0, max=vmax) for b in range(bsz)]
                out.append(torch.stack(col))
            return out
        grid = values[: bsz * nkeys].view(bsz, nkeys)
        return [grid[:, ki].clamp(0, self.vocab_sizes[k] - 1) for ki, k in enumerate(self.keys)]

    def forward(self, kjt) -> torch.Tensor:
        ids = self.decode_ids(kjt.values())
        return torch.cat([self.embeddings[k](ids[i]) for i, k in enumerate(self.keys)], dim=1)


class PortTower(nn.Module):
    def __init__(self, keys, vocab, emb_dim, dense_dim, hidden: Sequence[int], final_dim, dropout, decode):
        super().__init__()
        # construction order matters: it fixes RNG consumption so a seeded port has the
        # same initial weights as the seeded reference (base_tower.py:50-67)
        self.categorical_embedder = PortEmbedder(keys, vocab, emb_dim, decode)
        self.dense_projection = nn.Linear(dense_dim, hidden[0])
        layers: List[nn.Module] = []
        width = hidden[0] + len(keys) * emb_dim
        for h in hidden[1:]:
            layers += [nn.Linear(width, h), nn.ReLU(), nn.BatchNorm1d(h), nn.Dropout(dropout)]
            width = h
        layers.append(nn.Linear(width, final_dim))
        self.mlp = nn.Sequential(*layers)

    def forward(self, tower_input) -> torch.Tensor:
        proj = self.dense_projection(tower_input["dense"])
        cat = self.categorical_embedder(tower_input["kjt"])
        return F.normalize(self.mlp(torch.cat([proj, cat], dim=1)), p=2, dim=1)


class PortModel(nn.Module):
    def __init__(self, notice: PortTower, company: PortTower):
        super().__init__()
        self.notice_tower = notice
        self.company_tower = company

    def forward(self, notice_input, company_input):
        n = self.notice_tower(notice_input)
        c = self.company_tower(company_input)
        if n.size(0) != c.size(0):
            raise ValueError(f"batch size mismatch: {n.size(0)} vs {c.size(0)}")
        return n, c


class PortTrainTask(nn.Module):
    def __init__(self, model: PortModel, temperature=1.0, loss_type="cross_entropy", label_smoothing=0.0,
                 norm_check: bool = False):
        super().__init__()
        if loss_type not in ("cross_entropy", "cosine_embedding"):
            raise ValueError(f"Unsupported loss_type: {loss_type}")
        self.two_tower_model = model
        self.temperature = temperature
        self.loss_type = loss_type
        self.label_smoothing = label_smoothing
        self.norm_check = norm_check

    def similarity(self, n, c):
        s = torch.mm(n, c.t())
        return s / self.temperature if self.temperature != 1.0 else s

    def loss_fn(self, s):
        b = s.size(0)
        if self.loss_type == "cross_entropy":
            tgt = torch.arange(b, device=s.device)
            return 0.5 * (F.cross_entropy(s, tgt, label_smoothing=self.label_smoothing)
                          + F.cross_entropy(s.t(), tgt, label_smoothing=self.label_smoothing))
        # two_tower_train_task.py:136-158 (degenerate 1-D cosine_embedding_loss)
        off = ~torch.eye(b, dtype=torch.bool, device=s.device)
        sims = torch.cat([torch.diag(s), s[off]])
        tg = torch.cat([torch.ones(b, device=s.device), -torch.ones(b * b - b, device=s.device)])
        return F.cosine_embedding_loss(sims.unsqueeze(1), torch.ones_like(sims.unsqueeze(1)), tg)

    @staticmethod
    def metrics(s):
        b = s.size(0)
        acc = (s.argmax(dim=1) == torch.arange(b, device=s.device)).float().mean()
        diag = torch.diag(s)
        off = s[~torch.eye(b, dtype=torch.bool, device=s.device)]
        return {"accuracy": acc, "positive_similarity_mean": diag.mean(),
                "negative_similarity_mean": off.mean(), "similarity_gap": diag.mean() - off.mean()}

    def forward(self, batch, return_metrics=False):
        if batch["notice"]["dense"].size(0) != batch["company"]["dense"].size(0):
            raise ValueError("notice/company batch sizes differ")
        n, c = self.two_tower_model(batch["notice"], batch["company"])
        if self.norm_check:  # two_tower_model.py:81-87 — forces a sync in the reference
            torch.allclose(n.norm(dim=1), torch.ones(n.size(0)), atol=1e-4)
            torch.allclose(c.norm(dim=1), torch.ones(c.size(0)), atol=1e-4)
        s = self.similarity(n, c)
        loss = self.loss_fn(s)
        if not return_metrics:
            return loss
        return {"loss": loss, **self.metrics(s), "similarity_matrix": s.detach()}


def build_port_task(schema: dict, emb_dim=32, notice_dense=256, company_dense=128, hidden=(128, 64), final_dim=64,
                    dropout=0.0, temperature=1.0, loss_type="cross_entropy", label_smoothing=0.0, seed=0,
                    decode="vectorized") -> PortTrainTask:
    """schema = {"notice": [[key, vocab], ...], "company": [[key, vocab], ...]} (vocab already count+10)."""
    if seed is not None:
        torch.manual_seed(seed)
    nk = [k for k, _ in schema["notice"]]
    ck = [k for k, _ in schema["company"]]
    nt = PortTower(nk, dict(schema["notice"]), emb_dim, notice_dense, list(hidden), final_dim, dropout, decode)
    ct = PortTower(ck, dict(schema["company"]), emb_dim, company_dense, list(hidden), final_dim, dropout, decode)
    return PortTrainTask(PortModel(nt, ct), temperature, loss_type, label_smoothing)


def make_port_batch(schema, notice_ids, company_ids, notice_dense, company_dense):
    nk = [k for k, _ in schema["notice"]]
    ck = [k for k, _ in schema["company"]]
    return {"notice": {"dense": notice_dense, "kjt": PortKJT(nk, notice_ids.reshape(-1))},
            "company": {"dense": company_dense, "kjt": PortKJT(ck, company_ids.reshape(-1))}}


class PortTrainer:
    """zero_grad → forward(return_metrics=True) → backward → Adam.step → scheduler.step
    (scripts/train.py:231-242,319-330)."""

    def __init__(self, task: PortTrainTask, lr=1e-3, weight_decay=1e-5, warmup_steps: int = 0):
        self.task = task
        self.opt = torch.optim.Adam(task.parameters(), lr=lr, weight_decay=weight_decay)
        self.sched = None
        if warmup_steps and warmup_steps > 0:
            self.sched = torch.optim.lr_scheduler.LambdaLR(
                self.opt, lambda s: (s / warmup_steps) if s < warmup_steps else 1.0, last_epoch=-1)

    def step(self, batch):
        self.task.train()
        self.opt.zero_grad()
        out = self.task(batch, return_metrics=True)
        out["loss"].backward()
        self.opt.step()
        if self.sched is not None:
            self.sched.step()
        return out
