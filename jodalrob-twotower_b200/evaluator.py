"""TwoTowerEvaluator — same names, arguments and result keys as the reference's src/evaluation/evaluator.py, computed from one
integer per row (the rank of the positive pair) by hand-written kernels instead of topk + argsort + a Python loop over the rows.

  compute_recall_at_k(S, k), compute_mrr(S)       evaluator.py:20-71     -> tt_eval_ranks_from_matrix + tt_eval_rank_metrics
  evaluate_single_batch / evaluate_comprehensive  evaluator.py:119-209   -> model.eval_ranks(batch): tt_eval_ranks straight from the
                                                                            tower embeddings, the B x B matrix is never stored
Tie rule: a stable descending sort (ties keep column order); the reference's topk/argsort leave ties unspecified.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import native
from .native import TT_F32, ptr


def ranks_from_matrix(similarity_matrix: torch.Tensor, diag_offset: int = 0) -> torch.Tensor:
    """rank[i] (i32) = position of column diag_offset+i in a stable descending sort of row i."""
    s = similarity_matrix
    if not s.is_cuda:
        raise ValueError("similarity_matrix must live on the GPU (there is no CPU fallback)")
    if s.dtype != torch.float32 or s.stride(1) != 1:
        s = s.float().contiguous()
    rank = torch.empty(s.shape[0], dtype=torch.int32, device=s.device)
    native.lib().call("tt_eval_ranks_from_matrix", ptr(s), s.shape[0], s.shape[1], s.stride(0), diag_offset, ptr(rank),
                      torch.cuda.current_stream(s.device).cuda_stream)
    return rank


def ranks_from_embeddings(notice: torch.Tensor, company: torch.Tensor, inv_temp: float = 1.0, diag_offset: int = 0):
    """(rank i32 [Bn], positive logits f32 [Bn]) of S = notice @ company.T * inv_temp without materialising S."""
    n = notice.float().contiguous()
    c = company.float().contiguous()
    if not n.is_cuda or n.shape[1] != c.shape[1]:
        raise ValueError("embeddings must be CUDA tensors of equal width")
    rank = torch.empty(n.shape[0], dtype=torch.int32, device=n.device)
    diag = torch.empty(n.shape[0], dtype=torch.float32, device=n.device)
    native.lib().call("tt_eval_ranks", ptr(n), ptr(c), n.shape[0], c.shape[0], n.shape[1], TT_F32, float(inv_temp), diag_offset,
                      ptr(diag), ptr(rank), torch.cuda.current_stream(n.device).cuda_stream)
    return rank, diag


def rank_metrics(rank: torch.Tensor, ks: Sequence[int] = (5, 10)) -> torch.Tensor:
    """f32 [len(ks)+2] on the device: recall@k for every k (k is clamped to the row length like the reference's actual_k is a
    no-op for rank < B), MRR, top-1 accuracy."""
    if len(ks) > 8:
        raise ValueError("at most 8 values of k per call")
    d_ks = torch.tensor(list(ks), dtype=torch.int32, device=rank.device)
    out = torch.empty(len(ks) + 2, dtype=torch.float32, device=rank.device)
    native.lib().call("tt_eval_rank_metrics", ptr(rank), rank.numel(), ptr(d_ks), len(ks), ptr(out),
                      torch.cuda.current_stream(rank.device).cuda_stream)
    return out


class TwoTowerEvaluator:
    def __init__(self, device: str = "cuda:0"):
        self.device = device

    # ---- matrix-input entry points (evaluator.py:20-71)
    def compute_recall_at_k(self, similarity_matrix: torch.Tensor, k: int) -> torch.Tensor:
        return rank_metrics(ranks_from_matrix(similarity_matrix), (k,))[0]

    def compute_mrr(self, similarity_matrix: torch.Tensor) -> torch.Tensor:
        return rank_metrics(ranks_from_matrix(similarity_matrix), ())[0]

    def compute_comprehensive_metrics(self, similarity_matrix: torch.Tensor, basic_metrics: Dict) -> Dict[str, float]:
        r = rank_metrics(ranks_from_matrix(similarity_matrix), (5, 10)).tolist()
        return self._comprehensive(similarity_matrix.size(0), r[0], r[1], r[2], basic_metrics)

    @staticmethod
    def _comprehensive(batch_size: int, recall_5: float, recall_10: float, mrr: float, basic_metrics: Dict) -> Dict[str, float]:
        random_accuracy = 1.0 / batch_size
        random_recall_5 = min(5.0 / batch_size, 1.0)
        random_recall_10 = min(10.0 / batch_size, 1.0)
        return {
            "loss": basic_metrics.get("loss", 0.0), "accuracy": basic_metrics.get("accuracy", 0.0),
            "similarity_gap": basic_metrics.get("similarity_gap", 0.0),
            "positive_similarity_mean": basic_metrics.get("positive_similarity_mean", 0.0),
            "negative_similarity_mean": basic_metrics.get("negative_similarity_mean", 0.0),
            "recall@5": recall_5, "recall@10": recall_10, "mrr": mrr, "batch_size": batch_size,
            "random_accuracy": random_accuracy, "random_recall@5": random_recall_5, "random_recall@10": random_recall_10,
            "accuracy_improvement": basic_metrics.get("accuracy", 0.0) > random_accuracy,
            "recall@5_improvement": recall_5 > random_recall_5, "recall@10_improvement": recall_10 > random_recall_10,
        }

    # ---- model-input entry points (evaluator.py:119-209): fused, S is never stored
    def _batch_metrics(self, model, batch):
        """one eval forward -> (5 basic metrics + recall@5, recall@10, mrr) as ONE device tensor (a single D2H read)"""
        result = model(batch, return_metrics=True)
        if hasattr(model, "eval_ranks"):
            rank, _ = model.eval_ranks()
        else:  # any module with the reference's result dict
            rank = ranks_from_matrix(result["similarity_matrix"])
        basic = torch.stack([result[k].detach().float().reshape(()) for k in
                             ("loss", "accuracy", "similarity_gap", "positive_similarity_mean", "negative_similarity_mean")])
        return torch.cat([basic, rank_metrics(rank, (5, 10))[:3]]), rank.numel()

    def evaluate_single_batch(self, model, batch: Dict, verbose: bool = True) -> Dict[str, float]:
        model.eval()
        with torch.no_grad():
            vals, bsz = self._batch_metrics(model, batch)
            v = vals.tolist()
            basic = dict(zip(("loss", "accuracy", "similarity_gap", "positive_similarity_mean", "negative_similarity_mean"), v[:5]))
            metrics = self._comprehensive(bsz, v[5], v[6], v[7], basic)
        if verbose:
            self.print_single_batch_results(metrics)
        return metrics

    def evaluate_comprehensive(self, model, dataloader, verbose: bool = True) -> Dict[str, float]:
        model.eval()
        rows = []
        with torch.no_grad():
            for batch in dataloader:
                rows.append(self._batch_metrics(model, batch)[0])
        if not rows:
            raise ZeroDivisionError("empty dataloader")  # the reference divides by len(losses)
        avg = torch.stack(rows).double().mean(0).tolist()  # one D2H read for the whole evaluation
        out = {"loss": avg[0], "accuracy": avg[1], "recall@5": avg[5], "recall@10": avg[6], "mrr": avg[7], "similarity_gap": avg[2],
               "num_batches": len(rows)}
        if verbose:
            self.print_comprehensive_results(out)
        return out

    # ---- printing (plain-English equivalents of evaluator.py:211-272)
    def print_single_batch_results(self, metrics: Dict[str, float]):
        print(f"batch size: {int(metrics['batch_size'])}")
        print(f"Loss: {metrics['loss']:.4f}")
        print(f"Top-1 Accuracy: {metrics['accuracy']:.3f}")
        print(f"Recall@5: {metrics['recall@5']:.3f}")
        print(f"Recall@10: {metrics['recall@10']:.3f}")
        print(f"MRR: {metrics['mrr']:.3f}")
        print(f"Similarity Gap: {metrics['similarity_gap']:.3f}")
        print(f"random baselines: top-1 {metrics['random_accuracy']:.3f}, recall@5 {metrics['random_recall@5']:.3f}, "
              f"recall@10 {metrics['random_recall@10']:.3f}")

    def print_comprehensive_results(self, metrics: Dict[str, float]):
        print(f"test batches: {int(metrics['num_batches'])}")
        for k in ("loss", "accuracy", "recall@5", "recall@10", "mrr", "similarity_gap"):
            print(f"mean {k}: {metrics[k]:.4f}")
        self.print_performance_assessment(metrics)

    def print_performance_assessment(self, metrics: Dict[str, float]):
        a, r10, gap = metrics["accuracy"], metrics["recall@10"], metrics["similarity_gap"]
        print("top-1 accuracy: " + ("good (>0.3)" if a > 0.3 else "fair (0.15-0.3)" if a > 0.15 else "poor (<0.15)"))
        print("recall@10: " + ("practical (>0.6)" if r10 > 0.6 else "needs work (0.4-0.6)" if r10 > 0.4 else "insufficient (<0.4)"))
        print("similarity separation: " + ("fine (>0.5)" if gap > 0.5 else "needs work (<0.5)"))

    def demonstrate_predictions(self, model, batch: Dict, top_k: int = 10):
        model.eval()
        with torch.no_grad():
            p = model.predict_batch(batch, top_k=top_k)
            s = p["all_similarities"]
            print(f"batch size: {batch['notice']['dense'].shape[0]}")
            print(f"top-5 similarities of the first notice: {p['top_similarities'][0][:5]}")
            print(f"top-5 company indices of the first notice: {p['top_indices'][0][:5]}")
            print(f"similarity matrix: {tuple(s.shape)}, positives {torch.diag(s)[:5]}, first row {s[0].min():.3f} ~ {s[0].max():.3f}")
