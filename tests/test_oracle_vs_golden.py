"""Pins oracle/ref_port.py and oracle/two_tower_oracle.py to the golden vectors generated from the
unmodified reference (oracle/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_port, two_tower_oracle as orc
from tests.helpers import TINY_CASES, batch_inputs, load_golden, readme_schema, rel_err, sd_from


@pytest.mark.parametrize("case", TINY_CASES + ["cosine"])
def test_port_matches_reference_trajectory_bit_exact(case):
    torch.set_num_threads(1)
    d, cfg = load_golden(f"tiny_{case}.npz")
    task = ref_port.build_port_task(cfg["schema"], cfg["emb_dim"], cfg["notice_dense"], cfg["company_dense"], cfg["hidden"],
                                    cfg["final_dim"], 0.0, cfg["temperature"], cfg.get("loss_type", "cross_entropy"), cfg["label_smoothing"],
                                    seed=cfg["seed"])
    init = sd_from(d, "init.")
    # seeded construction reproduces the reference init bit-for-bit (same RNG consumption order)
    for k, v in task.state_dict().items():
        assert torch.equal(v, init[k]), k
    assert sum(p.numel() for p in task.parameters()) == cfg["n_params"]
    tr = ref_port.PortTrainer(task, cfg["lr"], cfg["weight_decay"], cfg["warmup_steps"])
    for i in range(3):
        nid, cid, nd, cd = batch_inputs(d, i)
        out = tr.step(ref_port.make_port_batch(cfg["schema"], nid, cid, nd, cd))
        assert np.float32(out["loss"].item()) == d[f"s{i}.loss"]
        assert torch.equal(out["similarity_matrix"], torch.from_numpy(d[f"s{i}.similarity_matrix"]))
        for m in ("accuracy", "positive_similarity_mean", "negative_similarity_mean", "similarity_gap"):
            assert np.float32(out[m].item()) == d[f"s{i}.{m}"]
    fin = sd_from(d, "final.")
    for k, v in task.state_dict().items():
        assert torch.equal(v, fin[k]), k


@pytest.mark.parametrize("case", TINY_CASES)
def test_port_loop_decode_equals_vectorized(case):
    d, cfg = load_golden(f"tiny_{case}.npz")
    nid, cid, nd, cd = batch_inputs(d, 0)
    outs = []
    for dec in ("loop", "vectorized"):
        task = ref_port.build_port_task(cfg["schema"], cfg["emb_dim"], cfg["notice_dense"], cfg["company_dense"], cfg["hidden"],
                                        cfg["final_dim"], 0.0, cfg["temperature"], seed=0, decode=dec)
        outs.append(task(ref_port.make_port_batch(cfg["schema"], nid, cid, nd, cd)))
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("case", TINY_CASES)
def test_explicit_oracle_matches_reference(case):
    torch.set_num_threads(1)
    d, cfg = load_golden(f"tiny_{case}.npz")
    st = orc.OracleStepper(sd_from(d, "init."), cfg["schema"], cfg["temperature"], cfg["label_smoothing"],
                           cfg["lr"], cfg["weight_decay"], cfg["warmup_steps"])
    for i in range(3):
        nid, cid, nd, cd = batch_inputs(d, i)
        loss, met, cache, grads, embs = st.step(nid.reshape(-1).numpy(), cid.reshape(-1).numpy(), nd, cd)
        assert abs(float(loss) - float(d[f"s{i}.loss"])) < 2e-6 * max(1.0, abs(float(d[f"s{i}.loss"])))
        assert rel_err(cache["s"], torch.from_numpy(d[f"s{i}.similarity_matrix"])) < 1e-5
        for m in ("accuracy", "positive_similarity_mean", "negative_similarity_mean", "similarity_gap"):
            assert abs(float(met[m]) - float(d[f"s{i}.{m}"])) < 1e-5
        if i == 0:
            assert rel_err(embs[0], torch.from_numpy(d["s0.notice_emb"])) < 1e-5
            assert rel_err(embs[1], torch.from_numpy(d["s0.company_emb"])) < 1e-5
            # hand-derived backward vs the reference's autograd, every parameter
            gexp = st_export_grads(st, grads)
            for k, g in gexp.items():
                ref = torch.from_numpy(d[f"grad0.{k}"])
                assert rel_err(g, ref) < 2e-5, k
    fin = sd_from(d, "final.")
    for k, v in st.export().items():
        if v.is_floating_point():
            assert rel_err(v, fin[k]) < 1e-5, k
        else:
            assert torch.equal(v, fin[k]), k


def st_export_grads(st, grads):
    out = {}
    for tw, keys in (("notice", st.nk), ("company", st.ck)):
        g, pre = grads[tw], f"two_tower_model.{tw}_tower."
        for k, t in zip(keys, g["emb"]):
            out[pre + f"categorical_embedder.embeddings.{k}.weight"] = t
        out[pre + "dense_projection.weight"], out[pre + "dense_projection.bias"] = g["Wd"], g["bd"]
        for bi, blk in enumerate(g["blocks"]):
            i = 4 * bi
            out[pre + f"mlp.{i}.weight"], out[pre + f"mlp.{i}.bias"] = blk["W"], blk["b"]
            out[pre + f"mlp.{i + 2}.weight"], out[pre + f"mlp.{i + 2}.bias"] = blk["gamma"], blk["beta"]
        i = st.P[tw]["final_index"]
        out[pre + f"mlp.{i}.weight"], out[pre + f"mlp.{i}.bias"] = g["Wf"], g["bf"]
    return out


def test_decode_clamp_matches_reference_clamping():
    d, cfg = load_golden("tiny_oob_temp.npz")
    nid, cid, _, _ = batch_inputs(d, 0)
    assert (nid < 0).any() and (nid >= torch.tensor([v for _, v in cfg["schema"]["notice"]])).any()
    vocab = [v for _, v in cfg["schema"]["notice"]]
    ids = orc.decode_clamp(nid.reshape(-1).numpy(), vocab)
    exp = torch.minimum(torch.clamp(nid, min=0), torch.tensor(vocab)[None] - 1).numpy()
    assert np.array_equal(ids, exp)
    # key-major is the transposed flattening of the same grid
    assert np.array_equal(orc.decode_clamp(nid.t().reshape(-1).numpy(), vocab, "key_major"), exp)
    # default vocab 1000 for a key with no category count (cat_embed.py:65-74), +10 margin otherwise
    sch = dict(cfg["schema"]["notice"])
    assert sch["nocount"] == 1000 and sch["flag_a"] == 12 and sch["org_code"] == 300


def test_readme_default_pins():
    torch.set_num_threads(1)
    d, cfg = load_golden("readme_default_b256.npz")
    assert cfg["n_params"] == 2204832 and cfg["n_params_train_py_config"] == 2890912  # train_results.csv:2-4
    schema = readme_schema()
    assert len(schema["notice"]) == 32 and len(schema["company"]) == 6
    assert sum(v for _, v in schema["notice"]) == 60024 and sum(v for _, v in schema["company"]) == 4117
    task = ref_port.build_port_task(schema, seed=0)
    assert sum(p.numel() for p in task.parameters()) == 2204832
    t2 = ref_port.build_port_task(schema, hidden=(512, 256), final_dim=128, seed=0)
    assert sum(p.numel() for p in t2.parameters()) == 2890912
    # loss ~ ln B at random init
    assert abs(float(d["s0.loss"]) - np.log(256)) < 0.05
    # unit-norm invariant (two_tower_model.py:84-87)
    assert np.allclose(np.linalg.norm(d["s0.notice_emb"], axis=1), 1.0, atol=1e-4)


def test_port_reproduces_readme_default_trajectory():
    import hashlib, json
    torch.set_num_threads(1)
    d, cfg = load_golden("readme_default_b256.npz")
    schema = readme_schema()
    task = ref_port.build_port_task(schema, seed=0)
    shas = json.loads(bytes(d["init_sha_json"]).decode())
    for k, v in task.state_dict().items():
        assert hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest() == shas[k], k
    gen = torch.Generator().manual_seed(cfg["input_seed"])
    tr = ref_port.PortTrainer(task, cfg["lr"], cfg["weight_decay"], cfg["warmup_steps"])
    for i in range(3):
        # same generator recipe as oracle/make_golden.py::_ids
        nid = torch.stack([torch.randint(0, v, (256,), generator=gen) for _, v in schema["notice"]], 1)
        cid = torch.stack([torch.randint(0, v, (256,), generator=gen) for _, v in schema["company"]], 1)
        nd, cd = torch.randn(256, 256, generator=gen), torch.randn(256, 128, generator=gen)
        assert np.array_equal(nid.numpy(), d[f"in{i}.notice_ids"]) and np.array_equal(cid.numpy(), d[f"in{i}.company_ids"])
        out = tr.step(ref_port.make_port_batch(schema, nid, cid, nd, cd))
        assert np.float32(out["loss"].item()) == d[f"s{i}.loss"]
    sd = task.state_dict()
    for k in sd:
        if "embeddings" in k:
            idx = torch.from_numpy(d[f"final_rows_idx.{k}"]).long()
            assert torch.equal(sd[k][idx], torch.from_numpy(d[f"final_rows.{k}"])), k
        else:
            assert torch.equal(sd[k], torch.from_numpy(np.array(d[f"final.{k}"]))), k


def test_rw_shard_and_sort_dedup_properties():
    rng = np.random.default_rng(0)
    vocab = [12, 300, 25, 1000, 51]
    ids = np.stack([rng.integers(0, v, 64) for v in vocab], 1)
    for world in (1, 2, 3, 8):
        owner, local = orc.rw_shard(ids, vocab, world)
        blk = np.array([orc.rw_block(v, world) for v in vocab])
        assert (owner >= 0).all() and (owner < world).all() and (local >= 0).all() and (local < blk[None]).all()
        assert np.array_equal(owner.astype(np.int64) * blk[None] + local, ids)
    rows = orc.arena_rows(ids, orc.row_offsets_of(vocab))
    sdp = orc.sort_dedup(rows)
    assert np.all(np.diff(sdp["sorted_rows"]) >= 0)
    assert np.array_equal(rows.reshape(-1)[sdp["positions"]], sdp["sorted_rows"])
    for s, e in zip(sdp["seg_starts"][:-1], sdp["seg_starts"][1:]):
        assert np.all(np.diff(sdp["positions"][s:e]) > 0)  # stable: positions ascend inside a segment
    assert np.array_equal(sdp["unique_rows"], np.unique(rows))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_retrieval_metric_oracle_matches_reference_evaluator(name):
    """oracle positive_ranks/recall_at_k/mrr vs the unmodified reference TwoTowerEvaluator (tests/golden/eval_metrics.npz,
    generated by oracle/make_golden.py eval_case from src/evaluation/evaluator.py:20-117)."""
    import json
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "eval_metrics.npz"))
    S = z[f"{name}.S"]
    for k in (1, 5, 10):
        assert abs(orc.recall_at_k(S, k) - float(z[f"{name}.recall@{k}"])) < 1e-7
    assert abs(orc.mrr(S) - float(z[f"{name}.mrr"])) < 1e-6
    cm = json.loads(bytes(z[f"{name}.comprehensive_json"]).decode())
    assert abs(cm["recall@10"] - orc.recall_at_k(S, 10)) < 1e-7 and cm["batch_size"] == S.shape[0]


def test_collate_oracle_matches_reference_collate_fn():
    """oracle collate_batch vs the batch the unmodified reference create_collate_fn produced (tests/golden/collate_batch.npz)."""
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "collate_batch.npz"))
    got = orc.collate_batch({"dense_projected": z["notice_store.dense"], "categorical": z["notice_store.cat"]},
                            {"dense_projected": z["company_store.dense"], "categorical": z["company_store.cat"]},
                            z["notice_idx"], z["company_idx"])
    for t in ("notice", "company"):
        assert np.array_equal(got[t]["dense"], z[f"batch.{t}.dense"])
        assert np.array_equal(got[t]["values"], z[f"batch.{t}.values"]) and got[t]["values"].dtype == np.int64
        assert np.array_equal(got[t]["lengths"], z[f"batch.{t}.lengths"])
    with pytest.raises(IndexError):
        orc.collate_batch({"dense_projected": z["notice_store.dense"], "categorical": z["notice_store.cat"]},
                          {"dense_projected": z["company_store.dense"], "categorical": z["company_store.cat"]}, [10 ** 6], [0])


def test_feature_projection_oracle_matches_reference_projector():
    """oracle feature_projection vs outputs of the unmodified reference FeatureProjector (tests/golden/projector.npz)."""
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "projector.npz"))
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    got = orc.feature_projection(sd, z["numeric"], {"title": z["text.title"], "body": z["text.body"]}, ["title", "body"])
    assert got.shape == z["dense_projected"].shape
    assert np.abs(got - z["dense_projected"]).max() <= 1e-5 * np.abs(z["dense_projected"]).max()
    only_text = orc.feature_projection(sd, None, {"body": z["text.body"]}, ["title", "body"])
    assert np.abs(only_text - z["proj.body"]).max() <= 1e-5 * np.abs(z["proj.body"]).max()


def test_pooled_oracle_consistency_with_torch_autograd_and_single_id_reference_case():
    """The L>1 (multi-hot) oracle has no reference counterpart (parity unpinned); what CAN be pinned on CPU:
    (1) with one id per bag it equals the reference-pinned decode_clamp + per-key gather; (2) its dense gradients equal torch
    autograd through F.embedding_bag(mode='sum') on clamped ids; (3) key-major and sample-major agree after the layout permutation."""
    g = torch.Generator().manual_seed(0)
    vocab, E, B = [7, 40, 300], 8, 13
    K = len(vocab)
    tables = [torch.randn(v, E, generator=g) for v in vocab]
    ids = torch.stack([torch.randint(-2, v + 3, (B,), generator=g) for v in vocab], 1)  # includes out-of-range ids
    unit = torch.arange(B * K + 1)
    out = orc.pooled_forward(tables, ids.reshape(-1), unit, B)
    dec = orc.decode_clamp(ids.reshape(-1).numpy(), vocab)  # [B, K], the reference-pinned clamp
    exp = torch.cat([tables[k][torch.from_numpy(dec[:, k]).long()] for k in range(K)], 1)
    assert torch.equal(out, exp)
    # jagged bags + gradients vs autograd
    lengths = torch.randint(0, 5, (B * K,), generator=g)
    offs = torch.zeros(B * K + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lengths, 0)
    vals = torch.cat([torch.randint(0, vocab[j % K], (int(lengths[j]),), generator=g) for j in range(B * K)])
    leaf = [t.clone().requires_grad_(True) for t in tables]
    parts = []
    for b in range(B):
        row = []
        for k in range(K):
            lo, hi = int(offs[b * K + k]), int(offs[b * K + k + 1])
            row.append(torch.nn.functional.embedding_bag(vals[lo:hi], leaf[k], torch.tensor([0]), mode="sum")[0] if hi > lo
                       else torch.zeros(E))
        parts.append(torch.cat(row))
    pooled = torch.stack(parts)
    assert torch.allclose(pooled, orc.pooled_forward(tables, vals, offs, B), atol=1e-6)
    grad = torch.randn(B, K * E, generator=g)
    pooled.backward(grad)
    gs = orc.pooled_dense_grads(vocab, E, vals, offs, B, grad)
    for k in range(K):
        assert torch.allclose(leaf[k].grad, gs[k], atol=1e-5), k
    touched = orc.pooled_touched(vocab, vals, offs, B)
    for k in range(K):
        assert torch.equal(touched[k], (gs[k].abs().sum(1) > 0) | touched[k])  # every row with a gradient is marked touched
    # key-major == sample-major after permuting the bags
    perm = torch.tensor([b * K + k for k in range(K) for b in range(B)])
    lens_km = lengths[perm]
    offs_km = torch.zeros(B * K + 1, dtype=torch.int64)
    offs_km[1:] = torch.cumsum(lens_km, 0)
    vals_km = torch.cat([vals[int(offs[j]):int(offs[j + 1])] for j in perm.tolist()]) if int(lengths.sum()) else vals
    assert torch.equal(orc.pooled_forward(tables, vals_km, offs_km, B, layout="key_major"), orc.pooled_forward(tables, vals, offs, B))


def test_oracle_integer_stage_properties():
    """Definition-level properties of the integer oracles the GPU tests compare against bit-exactly (randomised, CPU)."""
    rng = np.random.default_rng(12)
    for _ in range(20):
        B, n = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        S = rng.integers(-3, 4, (B, max(B, n))).astype(np.float32)  # many exact ties
        off = int(rng.integers(0, S.shape[1] - B + 1))
        r = orc.positive_ranks(S, off)
        for i in range(B):
            p, d = off + i, S[i, off + i]
            assert r[i] == int((S[i] > d).sum() + (S[i, :p] == d).sum())  # greater ones + equal ones before it (stable order)
        rows = rng.integers(0, 9, n).astype(np.int32)
        sd = orc.sort_dedup(rows)
        assert np.array_equal(sd["sorted_rows"], np.sort(rows, kind="stable"))
        assert np.array_equal(rows[sd["positions"]], sd["sorted_rows"])
        for a, b in zip(sd["seg_starts"][:-1], sd["seg_starts"][1:]):  # inside a row the positions ascend: fixed summation order
            assert (np.diff(sd["positions"][a:b]) > 0).all()
        vocab = [int(v) for v in rng.integers(1, 50, 4)]
        world = int(rng.integers(1, 6))
        ids = np.stack([rng.integers(0, v, 11) for v in vocab], 1)
        owner, local = orc.rw_shard(ids, vocab, world)
        blk = np.array([orc.rw_block(v, world) for v in vocab])
        assert ((owner >= 0) & (owner < world)).all() and (local >= 0).all() and (local < blk).all()
        assert np.array_equal(owner * blk + local, ids)  # the shard map is a bijection onto the ids
