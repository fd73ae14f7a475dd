"""TEST INFRASTRUCTURE ONLY — generate golden vectors from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

Writes small fixtures under tests/golden/ that travel to the GPU box:

  schema_readme_default.json  key order + vocab sizes derived from the reference's
                              meta/metadata.csv via its own classify_columns()
                              (data/column_classifier.py:67-130) and the +10 safety margin
                              (src/towers/cat_embed.py:26,70-77).  Derived data, not source.
  tiny_*.npz                  full-tensor trajectories on tests/golden/tiny_metadata.csv
                              (our own small schema run through the real reference):
                              init state_dict, 3 batches (with out-of-range ids), per-step
                              loss/metrics/similarity matrix, step-1 gradients of every
                              parameter, state_dict + Adam moments after 3 steps with the
                              LambdaLR warm-up of scripts/train.py:235-242.
  readme_default_b256.npz     README-default model (2,204,832 params) at B=256: init
                              checksums, inputs, 3-step loss trajectory, step-1 tower
                              embeddings, all dense params and sampled table rows after 3 steps.

The reference's own tests hold no golden vectors — these outputs of the
reference itself, run here, are what pins oracle/ref_port.py and oracle/two_tower_oracle.py.
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch

from oracle import ref_harness as rh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _vocab_of(task):
    nt = task.two_tower_model.notice_tower.categorical_embedder
    ct = task.two_tower_model.company_tower.categorical_embedder
    return ([[k, int(nt.vocab_sizes[k])] for k in nt.keys], [[k, int(ct.vocab_sizes[k])] for k in ct.keys])


def _ids(gen, vocab, bsz, oob_frac=0.0):
    cols = []
    for _, v in vocab:
        col = torch.randint(0, v, (bsz,), generator=gen)
        if oob_frac > 0:
            r = torch.rand(bsz, generator=gen)
            col = torch.where(r < oob_frac / 2, torch.full_like(col, -1), col)
            col = torch.where(r > 1 - oob_frac / 2, torch.full_like(col, v + 5), col)
        cols.append(col)
    return torch.stack(cols, 1).to(torch.int64)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def run_trajectory(task, nk, ck, batches, lr=1e-3, wd=1e-5, warmup_steps=2):
    """scripts/train.py:231-242,319-330 verbatim in spirit: Adam + LambdaLR + the loop body."""
    opt = torch.optim.Adam(task.parameters(), lr=lr, weight_decay=wd)
    sched = torch.optim.lr_scheduler.LambdaLR(
        opt, (lambda s: s / warmup_steps if s < warmup_steps else 1.0), last_epoch=-1)
    rec = {}
    task.train()
    for si, (nid, cid, nd, cd) in enumerate(batches):
        batch = rh.make_reference_batch(nk, ck, nid, cid, nd, cd)
        opt.zero_grad()
        hooks = []
        if si == 0:  # capture the tower outputs of this very forward (a 2nd forward would move BN running stats)
            hooks.append(task.two_tower_model.notice_tower.register_forward_hook(
                lambda m, i, o: rec.__setitem__("s0.notice_emb", o.detach().numpy().copy())))
            hooks.append(task.two_tower_model.company_tower.register_forward_hook(
                lambda m, i, o: rec.__setitem__("s0.company_emb", o.detach().numpy().copy())))
        with rh.silenced():
            out = task(batch, return_metrics=True)
        for h in hooks:
            h.remove()
        out["loss"].backward()
        if si == 0:
            for name, p in task.named_parameters():
                rec[f"grad0.{name}"] = p.grad.detach().numpy().copy()
        rec[f"s{si}.lr"] = np.float64(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
        for k in ("loss", "accuracy", "positive_similarity_mean", "negative_similarity_mean", "similarity_gap"):
            rec[f"s{si}.{k}"] = np.float32(out[k].item())
        rec[f"s{si}.similarity_matrix"] = out["similarity_matrix"].numpy().copy()
    return rec, opt


def tiny_case(name, hidden, final_dim, emb_dim, temperature, label_smoothing, oob, bsz=24, seed=0, loss_type="cross_entropy"):
    meta = os.path.join(GOLD, "tiny_metadata.csv")
    nd_dim, cd_dim = 12, 6
    task, nk, ck = rh.build_reference_task(
        categorical_embedding_dim=emb_dim, notice_dense_input_dim=nd_dim, company_dense_input_dim=cd_dim,
        tower_hidden_dims=hidden, final_embedding_dim=final_dim, dropout_rate=0.0,
        temperature=temperature, seed=seed, meta=meta, label_smoothing=label_smoothing, loss_type=loss_type)
    nv, cv = _vocab_of(task)
    out = {f"init.{k}": v.numpy().copy() for k, v in task.state_dict().items()}
    gen = torch.Generator().manual_seed(1234)
    batches = []
    for si in range(3):
        nid, cid = _ids(gen, nv, bsz, oob), _ids(gen, cv, bsz, oob)
        nd, cd = torch.randn(bsz, nd_dim, generator=gen), torch.randn(bsz, cd_dim, generator=gen)
        batches.append((nid, cid, nd, cd))
        out[f"in{si}.notice_ids"], out[f"in{si}.company_ids"] = nid.numpy(), cid.numpy()
        out[f"in{si}.notice_dense"], out[f"in{si}.company_dense"] = nd.numpy(), cd.numpy()
    rec, opt = run_trajectory(task, nk, ck, batches, warmup_steps=2)
    out.update(rec)
    for k, v in task.state_dict().items():
        out[f"final.{k}"] = v.numpy().copy()
    names = [n for n, _ in task.named_parameters()]
    for n, p in zip(names, opt.param_groups[0]["params"]):
        st = opt.state[p]
        out[f"adam_m.{n}"] = st["exp_avg"].numpy().copy()
        out[f"adam_v.{n}"] = st["exp_avg_sq"].numpy().copy()
    cfg = dict(schema={"notice": nv, "company": cv}, emb_dim=emb_dim, notice_dense=nd_dim, company_dense=cd_dim,
               hidden=list(hidden), final_dim=final_dim, temperature=temperature, label_smoothing=label_smoothing,
               lr=1e-3, weight_decay=1e-5, warmup_steps=2, bsz=bsz, seed=seed, loss_type=loss_type,
               n_params=int(sum(p.numel() for p in task.parameters())))
    out["config_json"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, f"tiny_{name}.npz"), **out)
    print(f"tiny_{name}: loss traj", [float(out[f's{i}.loss']) for i in range(3)], "params", cfg["n_params"])


def readme_default_case():
    task, nk, ck = rh.build_reference_task(dropout_rate=0.0, seed=0)
    nv, cv = _vocab_of(task)
    schema = {"notice": nv, "company": cv, "source": "derived from zoongahn/jodalroB-twoTower meta/metadata.csv "
              "via data/column_classifier.py:67-130; vocab = category count + 10 (src/towers/cat_embed.py:26,70-77)"}
    with open(os.path.join(GOLD, "schema_readme_default.json"), "w") as f:
        json.dump(schema, f, indent=1)
    bsz = 256
    out = {}
    init = task.state_dict()
    out["init_sha_json"] = np.frombuffer(json.dumps({k: sha(v) for k, v in init.items()}).encode(), dtype=np.uint8)
    gen = torch.Generator().manual_seed(1234)
    batches = []
    for si in range(3):
        nid, cid = _ids(gen, nv, bsz), _ids(gen, cv, bsz)
        nd, cd = torch.randn(bsz, 256, generator=gen), torch.randn(bsz, 128, generator=gen)
        batches.append((nid, cid, nd, cd))
        out[f"in{si}.notice_ids"], out[f"in{si}.company_ids"] = nid.numpy().astype(np.int32), cid.numpy().astype(np.int32)
        out[f"in{si}.notice_dense_sha"] = np.frombuffer(sha(nd).encode(), dtype=np.uint8)
        out[f"in{si}.company_dense_sha"] = np.frombuffer(sha(cd).encode(), dtype=np.uint8)
    rec, opt = run_trajectory(task, nk, ck, batches, warmup_steps=2)
    for k, v in rec.items():
        if k.startswith("grad0.") and "embeddings" in k:
            continue  # dense [V,E] table grads are large; touched rows are checked via the final rows
        if k.endswith("similarity_matrix"):
            continue
        out[k] = v
    final = task.state_dict()
    rows = {}
    g2 = torch.Generator().manual_seed(7)
    for k, v in final.items():
        if "embeddings" in k:
            n = v.shape[0]
            pick = torch.randperm(n, generator=g2)[: min(n, 16)].sort().values
            rows[k] = pick
            out[f"final_rows_idx.{k}"] = pick.numpy().astype(np.int32)
            out[f"final_rows.{k}"] = v[pick].numpy().copy()
            out[f"final_sum.{k}"] = np.float64(v.double().sum().item())
        else:
            out[f"final.{k}"] = v.numpy().copy()
    cfg = dict(emb_dim=32, notice_dense=256, company_dense=128, hidden=[128, 64], final_dim=64, temperature=1.0,
               label_smoothing=0.0, lr=1e-3, weight_decay=1e-5, warmup_steps=2, bsz=bsz, seed=0, input_seed=1234,
               n_params=int(sum(p.numel() for p in task.parameters())))
    # the scripts/train.py config (train_results.csv:2-4 pins 2,890,912 params)
    t2, _, _ = rh.build_reference_task(tower_hidden_dims=(512, 256), final_embedding_dim=128, seed=0)
    cfg["n_params_train_py_config"] = int(sum(p.numel() for p in t2.parameters()))
    out["config_json"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "readme_default_b256.npz"), **out)
    print("readme_default: loss traj", [float(out[f's{i}.loss']) for i in range(3)], cfg["n_params"], cfg["n_params_train_py_config"])


def eval_case():
    """Retrieval metrics of the UNMODIFIED reference evaluator (src/evaluation/evaluator.py:20-71) on fixed matrices."""
    rh.install_shims()
    from src.evaluation.evaluator import TwoTowerEvaluator
    ev = TwoTowerEvaluator(device="cpu")
    out = {}
    g = torch.Generator().manual_seed(77)
    for name, B, boost in (("a", 64, 1.0), ("b", 257, 2.5), ("c", 5, 0.0)):
        S = torch.randn(B, B, generator=g)
        S[torch.arange(B), torch.arange(B)] += boost
        out[f"{name}.S"] = S.numpy().copy()
        for k in (1, 5, 10):
            out[f"{name}.recall@{k}"] = np.float32(ev.compute_recall_at_k(S, k).item())
        out[f"{name}.mrr"] = np.float32(ev.compute_mrr(S).item())
        cm = ev.compute_comprehensive_metrics(S, {"loss": 1.25, "accuracy": 0.5})
        out[f"{name}.comprehensive_json"] = np.frombuffer(json.dumps({k: (bool(v) if isinstance(v, (bool, np.bool_)) else float(v))
                                                                       for k, v in cm.items()}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "eval_metrics.npz"), **out)
    print("eval_metrics:", {k: float(v) for k, v in out.items() if "recall@5" in k or "mrr" in k})


def projector_case():
    """Outputs of the UNMODIFIED reference FeatureProjector (src/torchrec_preprocess/feature_projector.py) and of the projection
    loop's concat order (feature_preprocessor.py:195-205) on fixed inputs."""
    rh.install_shims()
    from src.torchrec_preprocess.feature_projector import FeatureProjector
    torch.manual_seed(11)
    fp = FeatureProjector(num_dim=29, text_dim=768)
    g = torch.Generator().manual_seed(5)
    M = 41
    num = torch.randn(M, 29, generator=g)
    text = {"title": torch.randn(M, 768, generator=g), "body": torch.randn(M, 768, generator=g)}
    with torch.no_grad():
        dp, tp = fp(num, text)
        cat = torch.cat([dp, tp["title"], tp["body"]], dim=1)
    out = {f"sd.{k}": v.numpy().copy() for k, v in fp.state_dict().items()}
    out.update({"numeric": num.numpy(), "text.title": text["title"].numpy(), "text.body": text["body"].numpy(),
                "dense_proj": dp.numpy(), "proj.title": tp["title"].numpy(), "proj.body": tp["body"].numpy(), "dense_projected": cat.numpy()})
    np.savez_compressed(os.path.join(GOLD, "projector.npz"), **out)
    print("projector:", {k: v.shape for k, v in out.items() if not k.startswith("sd.")})


def collate_case():
    """Batch produced by the UNMODIFIED reference collate (create_collate_fn, unified_bid_data_loader.py:630-684, with its
    _build_batch_kjt :827-841) on small synthetic feature stores. sqlalchemy & co. are stubbed: the function only touches
    dataset.notice_store / company_store / schema.*.categorical."""
    import sys
    import types
    rh.install_shims()
    for name in ("sqlalchemy", "sqlalchemy.engine", "psycopg", "psycopg2", "tensorflow", "sentence_transformers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["sqlalchemy.engine"].Engine = object
    sys.modules["sqlalchemy"].engine = sys.modules["sqlalchemy.engine"]
    sys.modules["sqlalchemy"].create_engine = lambda *a, **k: None
    sys.modules["sqlalchemy"].text = lambda q: q
    from src.towers.pairs import unified_bid_data_loader as L
    rng = np.random.default_rng(5)
    nk = ["n_a", "n_b", "n_c", "n_d", "n_e"]
    ck = ["c_a", "c_b"]
    N, M, B = 97, 41, 19
    ds = types.SimpleNamespace(
        notice_store={"dense_projected": rng.standard_normal((N, 12)).astype(np.float32),
                      "categorical": rng.integers(-2, 500, (N, len(nk))).astype(np.int64)},
        company_store={"dense_projected": rng.standard_normal((M, 6)).astype(np.float32),
                       "categorical": rng.integers(0, 90, (M, len(ck))).astype(np.int64)},
        schema=types.SimpleNamespace(notice=types.SimpleNamespace(categorical=nk), company=types.SimpleNamespace(categorical=ck)))
    items = [{"notice_idx": int(rng.integers(0, N)), "company_idx": int(rng.integers(0, M)), "pair_info": {}} for _ in range(B)]
    batch = L.create_collate_fn(ds)(items)
    out = {"notice_store.dense": ds.notice_store["dense_projected"], "notice_store.cat": ds.notice_store["categorical"],
           "company_store.dense": ds.company_store["dense_projected"], "company_store.cat": ds.company_store["categorical"],
           "notice_idx": np.array([i["notice_idx"] for i in items], dtype=np.int64),
           "company_idx": np.array([i["company_idx"] for i in items], dtype=np.int64)}
    for t in ("notice", "company"):
        out[f"batch.{t}.dense"] = batch[t]["dense"].numpy().copy()
        out[f"batch.{t}.values"] = batch[t]["kjt"].values().numpy().copy()
        out[f"batch.{t}.lengths"] = batch[t]["kjt"].lengths().numpy().copy()
        out[f"batch.{t}.keys_json"] = np.frombuffer(json.dumps(list(batch[t]["kjt"].keys())).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, "collate_batch.npz"), **out)
    print("collate_batch:", {k: v.shape for k, v in out.items() if k.startswith("batch.") and "json" not in k})


def main():
    import sys
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order → reproducible fixtures
    if "projector" in sys.argv[1:] or len(sys.argv) == 1:
        projector_case()
        if len(sys.argv) > 1 and not {"cosine", "eval", "collate"} & set(sys.argv[1:]):
            return
    if "collate" in sys.argv[1:] or len(sys.argv) == 1:
        collate_case()
        if len(sys.argv) > 1 and not {"cosine", "eval"} & set(sys.argv[1:]):
            return
    if "eval" in sys.argv[1:] or len(sys.argv) == 1:
        eval_case()
        if len(sys.argv) > 1 and "cosine" not in sys.argv[1:]:
            return
    if "cosine" in sys.argv[1:] or len(sys.argv) == 1:
        # the non-default loss (two_tower_train_task.py:136-158): near-zero gradients, so the trajectory is driven by weight decay
        tiny_case("cosine", hidden=(16, 8), final_dim=8, emb_dim=8, temperature=0.5, label_smoothing=0.0, oob=0.0,
                  loss_type="cosine_embedding")
        if len(sys.argv) > 1:
            return
    tiny_case("base", hidden=(16, 8), final_dim=8, emb_dim=8, temperature=1.0, label_smoothing=0.0, oob=0.0)
    tiny_case("oob_temp", hidden=(16, 8), final_dim=8, emb_dim=8, temperature=0.5, label_smoothing=0.0, oob=0.1)
    tiny_case("deep_ls", hidden=(16, 12, 8), final_dim=4, emb_dim=4, temperature=0.25, label_smoothing=0.1, oob=0.05)
    tiny_case("nohidden", hidden=(16,), final_dim=8, emb_dim=8, temperature=1.0, label_smoothing=0.0, oob=0.0)
    readme_default_case()


if __name__ == "__main__":
    main()
