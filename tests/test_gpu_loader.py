"""GPU parity of the on-GPU batch assembly against the batch produced by the unmodified reference
collate function (tests/golden/collate_batch.npz) and against the numpy oracle at larger sizes. Byte/index work: bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as orc
from tests.helpers import GOLDEN, readme_schema

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_assembled_batch_equals_reference_collate_golden():
    from jodalrob_twotower_b200.loader import DeviceFeatureStore, GpuBatchAssembler, create_collate_fn
    z = np.load(os.path.join(GOLDEN, "collate_batch.npz"))
    nk = json.loads(bytes(z["batch.notice.keys_json"]).decode())
    ck = json.loads(bytes(z["batch.company.keys_json"]).decode())
    asm = GpuBatchAssembler(DeviceFeatureStore(z["notice_store.dense"], z["notice_store.cat"], nk, DEV),
                            DeviceFeatureStore(z["company_store.dense"], z["company_store.cat"], ck, DEV))
    items = [{"notice_idx": int(i), "company_idx": int(j), "pair_info": {}} for i, j in zip(z["notice_idx"], z["company_idx"])]
    batch = create_collate_fn(asm)(items)
    asm.check()
    for t, keys in (("notice", nk), ("company", ck)):
        assert batch[t]["dense"].is_cuda and batch[t]["dense"].dtype == torch.float32
        assert np.array_equal(batch[t]["dense"].cpu().numpy(), z[f"batch.{t}.dense"])
        assert np.array_equal(batch[t]["kjt"].values().cpu().numpy(), z[f"batch.{t}.values"])
        assert np.array_equal(batch[t]["kjt"].lengths().cpu().numpy(), z[f"batch.{t}.lengths"])
        assert batch[t]["kjt"].keys() == keys


@pytest.mark.parametrize("dense_dtype,cat_dtype", [(torch.float32, torch.int64), (torch.bfloat16, torch.int32), (torch.float16, torch.int64)])
@pytest.mark.parametrize("N,D,K,B", [(1, 4, 1, 1), (1000, 256, 32, 513), (50_000, 130, 6, 8192), (777, 7, 3, 100)])
def test_gather_vs_numpy_oracle(dense_dtype, cat_dtype, N, D, K, B):
    from jodalrob_twotower_b200.loader import DeviceFeatureStore
    rng = np.random.default_rng(N + D)
    dense = rng.standard_normal((N, D)).astype(np.float32)
    cat = rng.integers(-5, 100_000, (N, K)).astype(np.int64)
    st = DeviceFeatureStore(dense, cat, [f"k{i}" for i in range(K)], DEV, dense_dtype=dense_dtype, cat_dtype=cat_dtype)
    idx = rng.integers(0, N, B).astype(np.int64)
    # what the store holds after the (lossy for 16-bit) residency cast is the truth the gather must reproduce exactly
    held = {"dense_projected": torch.from_numpy(dense).to(dense_dtype).float().numpy(), "categorical": cat}
    exp = orc.collate_batch(held, held, idx, idx)["notice"]
    for layout in ("sample_major", "key_major"):
        d, v = st.gather(torch.from_numpy(idx).to(DEV), layout=layout)
        assert np.array_equal(d.cpu().numpy(), exp["dense"])
        ev = exp["values"] if layout == "sample_major" else exp["values"].reshape(B, K).T.reshape(-1)
        assert np.array_equal(v.cpu().numpy(), ev)


def test_out_of_range_index_raises_like_numpy_and_empty_batch():
    from jodalrob_twotower_b200.loader import DeviceFeatureStore, GpuBatchAssembler
    rng = np.random.default_rng(0)
    st = DeviceFeatureStore(rng.standard_normal((10, 8)).astype(np.float32), rng.integers(0, 9, (10, 2)), ["a", "b"], DEV)
    asm = GpuBatchAssembler(st, st)
    asm.assemble([0, 3], [9, 1], check=True)
    with pytest.raises(IndexError):
        asm.assemble([0, 10], [1, 2], check=True)
    with pytest.raises(IndexError):
        asm.assemble([0, 1], [-1, 2], check=True)
    with pytest.raises(ValueError):
        asm.assemble([0, 1, 2], [1, 2])  # the reference's batch-size mismatch contract (two_tower_train_task.py:64-67)
    b = asm.assemble([], [])
    assert b["notice"]["dense"].shape == (0, 8) and b["notice"]["kjt"].values().numel() == 0


def test_assemble_into_graphed_step_equals_host_collated_batch():
    """train through the captured graph fed by the on-GPU assembler == the same steps fed by host-collated batches, bit for bit"""
    import jodalrob_twotower_b200 as tt
    from jodalrob_twotower_b200.loader import DeviceFeatureStore, GpuBatchAssembler
    schema = readme_schema()
    nk, ck = [k for k, _ in schema["notice"]], [k for k, _ in schema["company"]]
    rng = np.random.default_rng(3)
    N, M, B = 5000, 3000, 512
    ns = {"dense_projected": rng.standard_normal((N, 256)).astype(np.float32),
          "categorical": np.stack([rng.integers(0, v, N) for _, v in schema["notice"]], 1).astype(np.int64)}
    cs = {"dense_projected": rng.standard_normal((M, 128)).astype(np.float32),
          "categorical": np.stack([rng.integers(0, v, M) for _, v in schema["company"]], 1).astype(np.int64)}

    def build():
        torch.manual_seed(0)
        task = tt.create_two_tower_train_task(nk, ck, categorical_embedding_dim=32, notice_dense_input_dim=256,
                                              company_dense_input_dim=128, tower_hidden_dims=[128, 64], final_embedding_dim=64,
                                              dropout_rate=0.1, temperature=1.0, device=DEV,
                                              vocab_sizes={"notice": dict(schema["notice"]), "company": dict(schema["company"])})
        tt.FusedAdam(task, lr=1e-3, weight_decay=1e-5)
        task.train()
        return task

    idxs = [(rng.integers(0, N, B), rng.integers(0, M, B)) for _ in range(3)]
    asm = GpuBatchAssembler(DeviceFeatureStore.from_dataset_store(ns, nk, DEV), DeviceFeatureStore.from_dataset_store(cs, ck, DEV))

    def host_batch(ni, ci):
        o = orc.collate_batch(ns, cs, ni, ci)
        mk = lambda keys, t: tt.KeyedJaggedTensor(keys, torch.from_numpy(o[t]["values"]).to(DEV))
        return {"notice": {"dense": torch.from_numpy(o["notice"]["dense"]).to(DEV), "kjt": mk(nk, "notice")},
                "company": {"dense": torch.from_numpy(o["company"]["dense"]).to(DEV), "kjt": mk(ck, "company")}}

    a, b = build(), build()
    ga = tt.GraphedFusedStep(a, host_batch(*idxs[0]))
    gb = tt.GraphedFusedStep(b, host_batch(*idxs[0]))
    for ni, ci in idxs:
        ra = ga(batch=host_batch(ni, ci)).clone()
        asm.assemble_into(gb, ni, ci)
        rb = gb().clone()
        assert torch.equal(ra, rb)
    asm.check()
    for (k, v), (_, v2) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, v2), k


def test_feature_projector_matches_reference_golden_and_feeds_the_store():
    """our FeatureProjector (same ctor / state_dict / init as the reference's) through tt_linear_forward
    vs outputs of the unmodified reference (tests/golden/projector.npz), fp32 tolerance 1e-5; project_store writes the
    concatenated layout straight into a device buffer that a DeviceFeatureStore can serve batches from."""
    from jodalrob_twotower_b200.loader import DeviceFeatureStore
    from jodalrob_twotower_b200.projector import FeatureProjector
    from tests.helpers import rel_err
    z = np.load(os.path.join(GOLDEN, "projector.npz"))
    torch.manual_seed(11)
    fp = FeatureProjector(num_dim=29, text_dim=768, device=DEV)
    sd = fp.state_dict()
    ref_sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    assert list(sd) == list(ref_sd)
    for k in sd:  # same module tree, same RNG consumption: seeded construction reproduces the reference init bit for bit
        assert torch.equal(sd[k].cpu(), ref_sd[k]), k
    num = torch.from_numpy(z["numeric"]).to(DEV)
    text = {"title": torch.from_numpy(z["text.title"]).to(DEV), "body": torch.from_numpy(z["text.body"]).to(DEV)}
    dp, tp = fp(num, text)
    assert rel_err(dp.cpu(), torch.from_numpy(z["dense_proj"])) < 1e-5
    assert rel_err(tp["title"].cpu(), torch.from_numpy(z["proj.title"])) < 1e-5
    assert rel_err(tp["body"].cpu(), torch.from_numpy(z["proj.body"])) < 1e-5
    # the store projection: numpy inputs, small batches (ragged last one), concat order numeric | title | body
    out = fp.project_store(z["numeric"], {"body": z["text.body"], "title": z["text.title"]}, ["title", "body"], batch_size=16)
    assert out.shape == (41, 384) and rel_err(out.cpu(), torch.from_numpy(z["dense_projected"])) < 1e-5
    exp = orc.feature_projection({k: v.numpy() for k, v in ref_sd.items()}, z["numeric"], {"title": z["text.title"], "body": z["text.body"]},
                                 ["title", "body"])
    assert np.abs(out.cpu().numpy() - exp).max() <= 1e-5 * np.abs(exp).max()
    assert fp.project_store(None, {}, None) is None and fp.project_store(z["numeric"][:0], None).shape == (0, 128)
    # ... and serves batches without ever leaving the GPU
    st = DeviceFeatureStore(out, np.zeros((41, 1), dtype=np.int64), ["k"], DEV)
    idx = torch.tensor([40, 0, 7, 7], device=DEV)
    d_, _ = st.gather(idx)
    assert torch.equal(d_, out[idx])
