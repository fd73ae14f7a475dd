"""TEST INFRASTRUCTURE ONLY — explicit-math CPU restatement of one reference train step.

Where ``oracle/ref_port.py`` restates the reference with torch modules + autograd, this
file writes every stage out as plain tensor arithmetic (forward, hand-derived backward,
Adam) so each CUDA kernel can be checked stage by stage, integer stages bit-exactly.
It is pinned by ``tests/test_oracle_vs_golden.py`` against fixtures generated from the
unmodified reference (``oracle/make_golden.py``).  The product never imports it.

Reference anchors (relative to /root/reference), see SURVEY.md Appendix A:
  A.1 decode/clamp        src/towers/cat_embed.py:98-121
  A.2 tower forward       src/towers/tower/base_tower.py:87-97,133-145
  A.3 logits + loss       src/towers/two_tower_train_task.py:106-134
  A.4 metrics             src/towers/two_tower_train_task.py:165-179
  A.6 optimizer           scripts/train.py:231-242 (torch.optim.Adam single-tensor math)
Multi-GPU row-wise shard map: no reference call site — TorchRec RW
convention restated here: block = ceil(V/W), owner = id // block, local = id - owner*block.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
NORM_EPS = 1e-12


# ----------------------------------------------------------------------------- integer stages
def decode_clamp(values: np.ndarray, vocab: Sequence[int], layout: str = "sample_major") -> np.ndarray:
    """values i64 [B*K] -> ids i64 [B,K]; cat_embed.py:98 (B = len // K), :109-110 (b*K + k), :115-117 (clamp)."""
    values = np.asarray(values, dtype=np.int64)
    k = len(vocab)
    b = values.shape[0] // k
    if layout == "sample_major":
        grid = values[: b * k].reshape(b, k)
    elif layout == "key_major":  # canonical TorchRec KJT order, behind an explicit flag
        grid = values[: b * k].reshape(k, b).T
    else:
        raise ValueError(layout)
    hi = np.asarray(vocab, dtype=np.int64)[None, :] - 1
    return np.minimum(np.maximum(grid, 0), hi)


def arena_rows(ids: np.ndarray, row_offsets: Sequence[int]) -> np.ndarray:
    """flat-arena row = row_offset[k] + id[b,k] (int32; the arena layout is ours)."""
    return (ids + np.asarray(row_offsets, dtype=np.int64)[None, : ids.shape[1]]).astype(np.int32)


def row_offsets_of(vocab: Sequence[int]) -> np.ndarray:
    off = np.zeros(len(vocab) + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.asarray(vocab, dtype=np.int64))
    return off


def rw_block(v: int, world: int) -> int:
    return -(-int(v) // int(world))


def rw_shard(ids: np.ndarray, vocab: Sequence[int], world: int) -> Tuple[np.ndarray, np.ndarray]:
    """(owner, local_row) per (b,k) under row-wise sharding."""
    blk = np.asarray([rw_block(v, world) for v in vocab], dtype=np.int64)[None, :]
    owner = ids // blk
    return owner.astype(np.int32), (ids - owner * blk).astype(np.int32)


def sort_dedup(rows: np.ndarray) -> Dict[str, np.ndarray]:
    """Stable sort of flat positions by arena row; unique rows, segment starts, position order.
    Determinism contract of the sparse backward: inside a segment, positions ascend."""
    flat = np.asarray(rows).reshape(-1).astype(np.int64)
    order = np.argsort(flat, kind="stable").astype(np.int32)
    srt = flat[order]
    head = np.ones(len(srt), dtype=bool)
    head[1:] = srt[1:] != srt[:-1]
    starts = np.flatnonzero(head).astype(np.int32)
    return {"sorted_rows": srt.astype(np.int32), "positions": order, "unique_rows": srt[head].astype(np.int32),
            "seg_starts": np.concatenate([starts, [len(srt)]]).astype(np.int32)}


# ----------------------------------------------------------------------------- parameters
def tower_params(sd: Dict[str, torch.Tensor], tower: str, keys: Sequence[str], dtype=torch.float32) -> dict:
    """Pull one tower's tensors out of a reference-named state_dict."""
    pre = f"two_tower_model.{tower}_tower."
    g = lambda n: sd[pre + n].detach().clone().to(dtype) if sd[pre + n].is_floating_point() else sd[pre + n].clone()
    p = {"emb": [g(f"categorical_embedder.embeddings.{k}.weight") for k in keys],
         "Wd": g("dense_projection.weight"), "bd": g("dense_projection.bias"), "blocks": []}
    i = 0
    while (pre + f"mlp.{i + 2}.running_mean") in sd:
        p["blocks"].append({"W": g(f"mlp.{i}.weight"), "b": g(f"mlp.{i}.bias"), "gamma": g(f"mlp.{i + 2}.weight"),
                            "beta": g(f"mlp.{i + 2}.bias"), "rm": g(f"mlp.{i + 2}.running_mean"),
                            "rv": g(f"mlp.{i + 2}.running_var"), "nbt": sd[pre + f"mlp.{i + 2}.num_batches_tracked"].clone()})
        i += 4
    p["Wf"], p["bf"] = g(f"mlp.{i}.weight"), g(f"mlp.{i}.bias")
    p["final_index"] = i
    return p


# ----------------------------------------------------------------------------- tower
def tower_forward(p: dict, dense: torch.Tensor, ids: torch.Tensor, train: bool = True,
                  dropout_p: float = 0.0, masks: Optional[List[torch.Tensor]] = None):
    """base_tower.py:133-145. ids are already clamped [B,K]. masks[i] (0/1 keep) feeds Dropout of block i."""
    cache = {"dense": dense, "ids": ids}
    proj = dense @ p["Wd"].t() + p["bd"]
    cat = torch.cat([p["emb"][k][ids[:, k]] for k in range(ids.shape[1])], dim=1)
    x = torch.cat([proj, cat], dim=1)
    cache["x0"] = x
    cache["blk"] = []
    bsz = x.shape[0]
    for bi, blk in enumerate(p["blocks"]):
        h = torch.relu(x @ blk["W"].t() + blk["b"])
        if train:
            mu = h.mean(0)
            var = ((h - mu) ** 2).mean(0)
            blk["rm"] = (1 - BN_MOMENTUM) * blk["rm"] + BN_MOMENTUM * mu
            blk["rv"] = (1 - BN_MOMENTUM) * blk["rv"] + BN_MOMENTUM * var * (bsz / max(bsz - 1, 1))
            blk["nbt"] = blk["nbt"] + 1
        else:
            mu, var = blk["rm"], blk["rv"]
        rstd = 1.0 / torch.sqrt(var + BN_EPS)
        xhat = (h - mu) * rstd
        y = xhat * blk["gamma"] + blk["beta"]
        keep = None
        if train and dropout_p > 0:
            keep = masks[bi].to(y.dtype)
            y = y * keep / (1.0 - dropout_p)
        cache["blk"].append({"xin": x, "h": h, "xhat": xhat, "rstd": rstd, "keep": keep})
        x = y
    z = x @ p["Wf"].t() + p["bf"]
    nrm = torch.clamp_min(z.norm(dim=1, keepdim=True), NORM_EPS)
    e = z / nrm
    cache.update(xlast=x, z=z, nrm=nrm, e=e, dropout_p=dropout_p, train=train)
    return e, cache


def tower_backward(p: dict, cache: dict, de: torch.Tensor) -> dict:
    """SURVEY.md A.5. Returns grads for every tower tensor; embedding grads are dense [V_k,E]."""
    g = {"blocks": []}
    e, nrm = cache["e"], cache["nrm"]
    dz = (de - e * (e * de).sum(1, keepdim=True)) / nrm
    g["Wf"], g["bf"] = dz.t() @ cache["xlast"], dz.sum(0)
    dx = dz @ p["Wf"]
    bsz = dx.shape[0]
    for blk, c in zip(reversed(p["blocks"]), reversed(cache["blk"])):
        dy = dx
        if c["keep"] is not None:
            dy = dy * c["keep"] / (1.0 - cache["dropout_p"])
        dgamma, dbeta = (dy * c["xhat"]).sum(0), dy.sum(0)
        dxhat = dy * blk["gamma"]
        if cache["train"]:
            dh = (c["rstd"] / bsz) * (bsz * dxhat - dxhat.sum(0) - c["xhat"] * (dxhat * c["xhat"]).sum(0))
        else:
            dh = dxhat * c["rstd"]
        dh = dh * (c["h"] > 0).to(dh.dtype)
        g["blocks"].insert(0, {"W": dh.t() @ c["xin"], "b": dh.sum(0), "gamma": dgamma, "beta": dbeta})
        dx = dh @ blk["W"]
    h0 = p["Wd"].shape[0]
    dproj, dcat = dx[:, :h0], dx[:, h0:]
    g["Wd"], g["bd"] = dproj.t() @ cache["dense"], dproj.sum(0)
    edim = p["emb"][0].shape[1]
    g["emb"] = []
    g["dcat"] = dcat
    for k, tab in enumerate(p["emb"]):
        gk = torch.zeros_like(tab)
        gk.index_add_(0, cache["ids"][:, k], dcat[:, k * edim:(k + 1) * edim])
        g["emb"].append(gk)
    return g


# ----------------------------------------------------------------------------- logits / loss / metrics
def simce_forward(n: torch.Tensor, c: torch.Tensor, temperature: float = 1.0, label_smoothing: float = 0.0):
    s = n @ c.t()
    if temperature != 1.0:
        s = s / temperature
    b = s.shape[0]
    lse_r = torch.logsumexp(s, dim=1)
    lse_c = torch.logsumexp(s, dim=0)
    diag = torch.diag(s)
    ls = label_smoothing
    l_row = -(1 - ls) * (diag - lse_r) - (ls / b) * (s.sum(1) - b * lse_r)
    l_col = -(1 - ls) * (diag - lse_c) - (ls / b) * (s.sum(0) - b * lse_c)
    loss = 0.5 * (l_row.mean() + l_col.mean())
    acc = (s.argmax(1) == torch.arange(b)).to(s.dtype).mean()
    pos = diag.mean()
    neg = (s.sum() - diag.sum()) / (b * b - b) if b > 1 else torch.tensor(float("nan"))
    met = {"accuracy": acc, "positive_similarity_mean": pos, "negative_similarity_mean": neg, "similarity_gap": pos - neg}
    return loss, met, {"s": s, "lse_r": lse_r, "lse_c": lse_c}


def simce_backward(n, c, cache, temperature: float = 1.0, label_smoothing: float = 0.0, dloss: float = 1.0):
    s, b = cache["s"], cache["s"].shape[0]
    pr = torch.exp(s - cache["lse_r"][:, None])
    pc = torch.exp(s - cache["lse_c"][None, :])
    tgt = (1 - label_smoothing) * torch.eye(b, dtype=s.dtype) + label_smoothing / b
    ds = (0.5 / b) * ((pr - tgt) + (pc - tgt)) * dloss
    if temperature != 1.0:
        ds = ds / temperature
    return ds @ c, ds.t() @ n


# ----------------------------------------------------------------------------- optimizer
def adam_update(w, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-5):
    """torch.optim.Adam (_single_tensor_adam, non-amsgrad, non-capturable) op order; in place.
    g may be None for 'no gradient this step' only in the *sparse* mode of the product; the dense
    reference semantics always pass a (possibly zero) gradient."""
    g = g + weight_decay * w if weight_decay != 0 else g
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    w.addcdiv_(m, denom, value=-step_size)


def warmup_lr(base_lr: float, sched_steps_so_far: int, warmup_steps: int) -> float:
    """LambdaLR of scripts/train.py:235-242: factor s/warmup for s<warmup else 1 (0 at the first step)."""
    if warmup_steps <= 0:
        return base_lr
    return base_lr * (sched_steps_so_far / warmup_steps if sched_steps_so_far < warmup_steps else 1.0)


# ----------------------------------------------------------------------------- whole step
class OracleStepper:
    """Holds reference-named tensors; step() = forward, backward, dense Adam on every tensor."""

    def __init__(self, sd: Dict[str, torch.Tensor], schema: dict, temperature=1.0, label_smoothing=0.0,
                 lr=1e-3, weight_decay=1e-5, warmup_steps=0, dtype=torch.float32):
        self.nk = [k for k, _ in schema["notice"]]
        self.ck = [k for k, _ in schema["company"]]
        self.nv = [v for _, v in schema["notice"]]
        self.cv = [v for _, v in schema["company"]]
        self.P = {"notice": tower_params(sd, "notice", self.nk, dtype), "company": tower_params(sd, "company", self.ck, dtype)}
        self.T, self.ls, self.lr, self.wd, self.warm = temperature, label_smoothing, lr, weight_decay, warmup_steps
        self.t = 0
        self.state: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def _leaves(self, tree_p, tree_g):
        for tw in ("notice", "company"):
            p, g = tree_p[tw], tree_g[tw]
            for a, b in zip(p["emb"], g["emb"]):
                yield a, b
            yield p["Wd"], g["Wd"]
            yield p["bd"], g["bd"]
            for pb, gb in zip(p["blocks"], g["blocks"]):
                for n in ("W", "b", "gamma", "beta"):
                    yield pb[n], gb[n]
            yield p["Wf"], g["Wf"]
            yield p["bf"], g["bf"]

    def forward_backward(self, notice_values, company_values, notice_dense, company_dense, train=True):
        nid = torch.from_numpy(decode_clamp(np.asarray(notice_values), self.nv))
        cid = torch.from_numpy(decode_clamp(np.asarray(company_values), self.cv))
        dt = self.P["notice"]["Wd"].dtype
        en, cn = tower_forward(self.P["notice"], notice_dense.to(dt), nid, train)
        ec, cc = tower_forward(self.P["company"], company_dense.to(dt), cid, train)
        loss, met, cache = simce_forward(en, ec, self.T, self.ls)
        dn, dc = simce_backward(en, ec, cache, self.T, self.ls)
        grads = {"notice": tower_backward(self.P["notice"], cn, dn), "company": tower_backward(self.P["company"], cc, dc)}
        return loss, met, cache, grads, (en, ec), (dn, dc)

    def step(self, notice_values, company_values, notice_dense, company_dense):
        loss, met, cache, grads, embs, _ = self.forward_backward(notice_values, company_values, notice_dense, company_dense)
        lr = warmup_lr(self.lr, self.t, self.warm)
        self.t += 1
        for w, g in self._leaves(self.P, grads):
            if id(w) not in self.state:
                self.state[id(w)] = (torch.zeros_like(w), torch.zeros_like(w))
            m, v = self.state[id(w)]
            adam_update(w, g, m, v, self.t, lr, weight_decay=self.wd)
        return loss, met, cache, grads, embs

    def export(self) -> Dict[str, torch.Tensor]:
        out = {}
        for tw, keys in (("notice", self.nk), ("company", self.ck)):
            p, pre = self.P[tw], f"two_tower_model.{tw}_tower."
            for k, t in zip(keys, p["emb"]):
                out[pre + f"categorical_embedder.embeddings.{k}.weight"] = t
            out[pre + "dense_projection.weight"], out[pre + "dense_projection.bias"] = p["Wd"], p["bd"]
            for bi, blk in enumerate(p["blocks"]):
                i = 4 * bi
                out[pre + f"mlp.{i}.weight"], out[pre + f"mlp.{i}.bias"] = blk["W"], blk["b"]
                out[pre + f"mlp.{i + 2}.weight"], out[pre + f"mlp.{i + 2}.bias"] = blk["gamma"], blk["beta"]
                out[pre + f"mlp.{i + 2}.running_mean"], out[pre + f"mlp.{i + 2}.running_var"] = blk["rm"], blk["rv"]
                out[pre + f"mlp.{i + 2}.num_batches_tracked"] = blk["nbt"]
            i = p["final_index"]
            out[pre + f"mlp.{i}.weight"], out[pre + f"mlp.{i}.bias"] = p["Wf"], p["bf"]
        return out


# ----------------------------------------------------------------------------- pooled (multi-hot) tables, L >= 0 ids per bag
def pooled_forward(tables: Sequence[torch.Tensor], values: torch.Tensor, offsets: torch.Tensor, B: int,
                   layout: str = "sample_major") -> torch.Tensor:
    """Sum-pooled EmbeddingBagCollection forward -> [B, K*E]. The reference has one id per bag
    (unified_bid_data_loader.py:835); L>1 is the BASELINE config-5 generalisation whose oracle is
    torch embedding_bag(mode="sum"). Bag j = b*K+k (sample-major) or k*B+b (key-major);
    ids clamp to [0, V_k-1] like cat_embed.py:115-117."""
    K, E = len(tables), tables[0].shape[1]
    out = torch.zeros(B, K * E, dtype=tables[0].dtype)
    for b in range(B):
        for k in range(K):
            j = b * K + k if layout == "sample_major" else k * B + b
            lo, hi = int(offsets[j]), int(offsets[j + 1])
            if hi > lo:
                ids = values[lo:hi].clamp(0, tables[k].shape[0] - 1)
                out[b, k * E:(k + 1) * E] = torch.nn.functional.embedding_bag(ids, tables[k], torch.tensor([0]), mode="sum")[0]
    return out


def pooled_dense_grads(vocab: Sequence[int], E: int, values: torch.Tensor, offsets: torch.Tensor, B: int, grad: torch.Tensor,
                       layout: str = "sample_major") -> List[torch.Tensor]:
    """Dense [V_k, E] gradient of every table (what autograd's embedding_dense_backward produces), accumulated in
    ascending value position, the order the deterministic sort-dedup fixes."""
    K = len(vocab)
    gs = [torch.zeros(v, E, dtype=grad.dtype) for v in vocab]
    for b in range(B):
        for k in range(K):
            j = b * K + k if layout == "sample_major" else k * B + b
            lo, hi = int(offsets[j]), int(offsets[j + 1])
            if hi > lo:
                ids = values[lo:hi].clamp(0, vocab[k] - 1)
                gs[k].index_add_(0, ids, grad[b, k * E:(k + 1) * E].expand(hi - lo, E))
    return gs


def pooled_touched(vocab: Sequence[int], values: torch.Tensor, offsets: torch.Tensor, B: int, layout: str = "sample_major"):
    K = len(vocab)
    t = [torch.zeros(v, dtype=torch.bool) for v in vocab]
    for b in range(B):
        for k in range(K):
            j = b * K + k if layout == "sample_major" else k * B + b
            lo, hi = int(offsets[j]), int(offsets[j + 1])
            if hi > lo:
                t[k][values[lo:hi].clamp(0, vocab[k] - 1)] = True
    return t


# ----------------------------------------------------------------------------- retrieval metrics (evaluator.py:20-71)
def positive_ranks(S: np.ndarray, diag_offset: int = 0) -> np.ndarray:
    """0-based position of column diag_offset+i in a STABLE descending sort of row i — what
    `(argsort(S, descending=True)[i] == i).nonzero()` (evaluator.py:56-66) yields when ties keep column order."""
    S = np.asarray(S)
    order = np.argsort(-S, axis=1, kind="stable")
    pos = np.arange(S.shape[0]) + diag_offset
    return np.argmax(order == pos[:, None], axis=1).astype(np.int32)


def recall_at_k(S: np.ndarray, k: int) -> float:
    """evaluator.py:20-43: mean over rows of [positive among the top-k columns]."""
    return float((positive_ranks(S) < min(k, S.shape[1])).astype(np.float32).mean())


def mrr(S: np.ndarray) -> float:
    """evaluator.py:45-71: mean of 1/(rank+1) (python floats, then a float32 tensor mean)."""
    return float(np.mean((1.0 / (positive_ranks(S).astype(np.float64) + 1.0)).astype(np.float32)))


# ----------------------------------------------------------------------------- batch assembly (collate)
def collate_batch(notice_store: dict, company_store: dict, notice_idx, company_idx) -> dict:
    """create_collate_fn + _build_batch_kjt (unified_bid_data_loader.py:630-684,827-841) restated with numpy:
    dense = store['dense_projected'][idx] as float32, values = store['categorical'][idx] flattened SAMPLE-major as int64,
    lengths = ones. numpy raises IndexError for an out-of-range index, like the reference."""
    out = {}
    for name, st, idx in (("notice", notice_store, notice_idx), ("company", company_store, company_idx)):
        idx = np.asarray(idx, dtype=np.int64)
        cat = np.asarray(st["categorical"])[idx]
        out[name] = {"dense": np.asarray(st["dense_projected"])[idx].astype(np.float32),
                     "values": cat.reshape(-1).astype(np.int64), "lengths": np.ones(cat.size, dtype=np.int64)}
    return out


# ----------------------------------------------------------------------------- feature pre-projection
def feature_projection(sd: Dict[str, np.ndarray], numeric: Optional[np.ndarray], text: Dict[str, np.ndarray],
                       text_cols: Sequence[str], dtype=np.float64) -> np.ndarray:
    """FeatureProjector (src/torchrec_preprocess/feature_projector.py:4-28) + the concat order of the store projection
    (feature_preprocessor.py:195-205): [Linear2(relu(Linear1(numeric))) | the same text MLP on every text column in order]."""
    def mlp(x, pre):
        h = np.maximum(x.astype(dtype) @ sd[f"{pre}.0.weight"].astype(dtype).T + sd[f"{pre}.0.bias"].astype(dtype), 0)
        return h @ sd[f"{pre}.2.weight"].astype(dtype).T + sd[f"{pre}.2.bias"].astype(dtype)
    parts = ([mlp(numeric, "num_proj")] if numeric is not None else []) + [mlp(text[c], "text_proj") for c in text_cols if c in text]
    return np.concatenate(parts, axis=1)
