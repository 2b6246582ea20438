"""CPU restatement of the rank metrics. Test infrastructure -- see ``oracle/__init__.py``.

* ``ref_rank``        = ``Rank.update``          ``aligner/metrics.py:16-19`` (literal argsort + where)
* ``ref_median_rank`` = ``MedianRank.compute``   ``aligner/metrics.py:33-36`` (torch lower median + 1)
* ``ref_recall_at_k`` = ``torchmetrics.Recall(top_k=k)`` / ``Accuracy(top_k=k)`` with the 0.9 defaults
  (multiclass, ``average="micro"``) [3P, ``aligner/text_video_retrieval.py:21``,
  ``aligner/video_text_classification.py:61``]: the fraction of rows whose target column is among the k largest.
* ``ref_stable_rank`` = the documented tie rule of the CUDA path, in numpy-style integer arithmetic:
  ``rank_i = #{j: s_ij > s_it} + #{j < t: s_ij == s_it}`` (what a stable descending sort gives; the reference's
  unstable argsort leaves ties implementation-defined).
"""
from __future__ import annotations

from typing import Dict

import torch


def ref_rank(predictions: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    sorted_predicted_positions = predictions.argsort(dim=1, descending=True)
    return torch.where(sorted_predicted_positions == target.unsqueeze(-1))[1]


def ref_stable_rank(predictions: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    target = target.to(torch.int64)
    t_score = predictions.gather(1, target.unsqueeze(-1))
    greater = (predictions > t_score).sum(dim=1)
    cols = torch.arange(predictions.shape[1]).unsqueeze(0)
    tied_before = ((predictions == t_score) & (cols < target.unsqueeze(-1))).sum(dim=1)
    return greater + tied_before


def ref_median_rank(ranks: torch.Tensor) -> torch.Tensor:
    return ranks.median() + 1


def ref_recall_at_k(predictions: torch.Tensor, target: torch.Tensor, k: int = 1) -> torch.Tensor:
    topk = predictions.topk(k, dim=1).indices
    hits = (topk == target.unsqueeze(-1)).any(dim=1)
    return hits.sum().to(torch.float32) / predictions.shape[0]


ref_accuracy_at_k = ref_recall_at_k  # micro multiclass accuracy == micro multiclass recall


def ref_retrieval_metrics(scores: torch.Tensor) -> Dict[str, torch.Tensor]:
    """``TextVideoRetrievalLightningModule._validate_dataset`` (``aligner/text_video_retrieval.py:67-83``):
    target of row i is column i."""
    target = torch.arange(scores.shape[-1])
    ranks = ref_rank(scores, target)
    return {"r1": ref_recall_at_k(scores, target, 1), "r5": ref_recall_at_k(scores, target, 5),
            "r10": ref_recall_at_k(scores, target, 10), "mr": ref_median_rank(ranks), "rank": ranks}
