"""GPU parity of the fused retrieval metrics against the oracle restatement of
src/evaluation/evaluator.py:20-71 and against torch's own topk/argsort on the same matrix. Ranks are integers: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as orc
from tests.helpers import batch_inputs, load_golden, sd_from
from tests.test_gpu_parity import make_batch, make_task

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def unit_rows(B, F, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(B, F, generator=g), dim=1)


@pytest.mark.parametrize("Bx,By,off,F,T", [(1, 1, 0, 8, 1.0), (24, 24, 0, 8, 0.5), (200, 200, 0, 64, 1.0), (1000, 1000, 0, 64, 0.05),
                                           (130, 700, 300, 16, 1.0), (4096, 4096, 0, 64, 1.0)])
def test_fused_ranks_bit_exact_vs_materialised_matrix(Bx, By, off, F, T):
    from jodalrob_twotower_b200 import evaluator as ev, native, ops
    x, y = unit_rows(Bx, F, 1).to(DEV), unit_rows(By, F, 2).to(DEV)
    if By > 3:  # exact ties: duplicate company rows, before and after the positive's column
        y[1] = y[By - 1]
        y[min(By - 2, off + 2)] = y[off]
    rank, diag = ev.ranks_from_embeddings(x, y, 1.0 / T, off)
    S = torch.empty(Bx, By, device=DEV)
    native.lib().call("tt_similarity_matrix", x.data_ptr(), y.data_ptr(), Bx, By, F, native.TT_F32, 1.0 / T, S.data_ptr(), By,
                      torch.cuda.current_stream().cuda_stream)
    exp = orc.positive_ranks(S.cpu().numpy(), off)
    assert np.array_equal(rank.cpu().numpy(), exp)                                   # fused == oracle on the same logits
    assert np.array_equal(ev.ranks_from_matrix(S, off).cpu().numpy(), exp)           # matrix entry point
    assert torch.equal(diag, S[torch.arange(Bx), torch.arange(Bx) + off])            # bit-identical positive logits
    # against the reference's own formulation on an fp64 matrix (no ties left to chance: only where sims are well separated)
    S64 = (x.double() @ y.double().t() / T).cpu().numpy()
    exp64 = orc.positive_ranks(S64, off)
    assert (rank.cpu().numpy() != exp64).mean() <= 0.01  # fp32 rounding can swap near-equal neighbours, nothing more


def test_recall_and_mrr_match_reference_formulas():
    from jodalrob_twotower_b200 import TwoTowerEvaluator
    B = 513
    g = torch.Generator().manual_seed(5)
    S = torch.randn(B, B, generator=g)
    S[torch.arange(B), torch.arange(B)] += 1.5  # positives mostly near the top
    Sd = S.to(DEV)
    e = TwoTowerEvaluator(DEV)
    for k in (1, 5, 10, 600):
        # evaluator.py:33-41 verbatim
        topk = torch.topk(S, k=min(k, B), dim=1).indices
        ref = (topk == torch.arange(B).unsqueeze(1)).any(dim=1).float().mean().item()
        assert abs(e.compute_recall_at_k(Sd, k).item() - ref) < 1e-7
        assert abs(orc.recall_at_k(S.numpy(), k) - ref) < 1e-7
    # evaluator.py:56-70 verbatim
    srt = torch.argsort(S, dim=1, descending=True)
    rr = [1.0 / ((srt[i] == i).nonzero().item() + 1) for i in range(B)]
    ref = torch.tensor(rr).mean().item()
    assert abs(e.compute_mrr(Sd).item() - ref) < 1e-6
    assert abs(orc.mrr(S.numpy()) - ref) < 1e-6
    m = e.compute_comprehensive_metrics(Sd, {"loss": 1.0, "accuracy": 0.5})
    assert m["batch_size"] == B and abs(m["recall@10"] - e.compute_recall_at_k(Sd, 10).item()) < 1e-7
    assert m["random_recall@5"] == min(5.0 / B, 1.0) and m["accuracy_improvement"] is True


def test_evaluator_on_model_never_needs_the_matrix():
    from jodalrob_twotower_b200 import TwoTowerEvaluator
    d, cfg = load_golden("tiny_oob_temp.npz")
    task = make_task(cfg, cfg["schema"])
    task.load_state_dict(sd_from(d, "init."))
    batches = [make_batch(cfg["schema"], *batch_inputs(d, i)) for i in range(3)]
    e = TwoTowerEvaluator(DEV)
    single = e.evaluate_single_batch(task, batches[0], verbose=False)
    task.eval()
    with torch.no_grad():
        out = task(batches[0], return_metrics=True)
        S = out["similarity_matrix"].cpu().numpy()
    assert abs(single["recall@5"] - orc.recall_at_k(S, 5)) < 1e-7 and abs(single["recall@10"] - orc.recall_at_k(S, 10)) < 1e-7
    assert abs(single["mrr"] - orc.mrr(S)) < 1e-6
    assert abs(single["loss"] - out["loss"].item()) < 1e-7 and single["batch_size"] == cfg["bsz"]
    assert abs(single["accuracy"] - float((orc.positive_ranks(S) == 0).mean())) < 1e-7  # top-1 == rank 0
    allm = e.evaluate_comprehensive(task, batches, verbose=False)
    per = [e.evaluate_single_batch(task, b, verbose=False) for b in batches]
    for k in ("loss", "accuracy", "recall@5", "recall@10", "mrr", "similarity_gap"):
        assert abs(allm[k] - sum(p[k] for p in per) / 3) < 1e-6, k
    assert allm["num_batches"] == 3


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_evaluator_matches_reference_golden(name):
    """the shipped TwoTowerEvaluator against outputs of the unmodified reference evaluator on the same matrices"""
    import json
    import os
    from jodalrob_twotower_b200 import TwoTowerEvaluator
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "eval_metrics.npz"))
    S = torch.from_numpy(z[f"{name}.S"]).to(DEV)
    e = TwoTowerEvaluator(DEV)
    for k in (1, 5, 10):
        assert abs(e.compute_recall_at_k(S, k).item() - float(z[f"{name}.recall@{k}"])) < 1e-7
    assert abs(e.compute_mrr(S).item() - float(z[f"{name}.mrr"])) < 1e-6
    ref = json.loads(bytes(z[f"{name}.comprehensive_json"]).decode())
    got = e.compute_comprehensive_metrics(S, {"loss": 1.25, "accuracy": 0.5})
    assert set(got) == set(ref)
    for k, v in ref.items():
        assert abs(float(got[k]) - float(v)) < 1e-6, k
