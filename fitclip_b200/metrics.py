"""Rank metrics with the ``torchmetrics.Metric`` calling protocol the reference uses (``metric(preds, target)`` to
update, ``compute()``, ``reset()``): ``Rank`` / ``MeanRank`` / ``MedianRank`` (``aligner/metrics.py:6-36``) and the
micro top-k ``Recall`` / ``Accuracy`` of torchmetrics 0.9 (``aligner/text_video_retrieval.py:21``,
``aligner/video_text_classification.py:61``).  State is a list of int64 rank tensors ("cat" reduction, ``metrics.py:13``);
every metric derives from the ranks, so one pass over the score matrix serves all of them.

Tie rule (the reference's unstable argsort leaves it undefined): a tied column outranks the
target only if its index is lower -- what a stable descending sort yields."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops


class Rank:
    is_differentiable = False
    higher_is_better = False
    full_state_update = False

    def __init__(self, process_group: Optional[dist.ProcessGroup] = None) -> None:
        self.ranks: List[torch.Tensor] = []
        self.num_candidates = 0
        self.process_group = process_group

    def update(self, predictions: torch.Tensor, target: torch.Tensor) -> None:
        predictions = predictions if predictions.dtype == torch.float32 else predictions.float()
        self.ranks.append(ops.rank_from_scores(predictions, target))
        self.num_candidates = max(self.num_candidates, predictions.shape[1])

    def update_from_ranks(self, ranks: torch.Tensor, num_candidates: int) -> None:
        """Feed ranks computed elsewhere (the fused similarity+count kernel never builds ``predictions``)."""
        self.ranks.append(ranks.to(torch.int64))
        self.num_candidates = max(self.num_candidates, num_candidates)

    def __call__(self, predictions: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        self.update(predictions, target)
        return self._compute_from(self.ranks[-1])

    def _all_ranks(self) -> torch.Tensor:
        if not self.ranks:
            raise RuntimeError("compute() called before update()")
        ranks = torch.cat(self.ranks)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            # dist_reduce_fx="cat" (metrics.py:13); shards may be uneven: one size gather (one host read), one padded
            # data gather
            world = dist.get_world_size(self.process_group)
            n = torch.tensor([ranks.numel()], device=ranks.device, dtype=torch.int64)
            sizes_t = torch.empty(world, device=ranks.device, dtype=torch.int64)
            dist.all_gather_into_tensor(sizes_t, n, group=self.process_group)
            sizes = sizes_t.tolist()
            cap = max(sizes)
            padded = torch.zeros(cap, dtype=ranks.dtype, device=ranks.device)
            padded[:ranks.numel()] = ranks
            parts = torch.empty(world * cap, dtype=ranks.dtype, device=ranks.device)
            dist.all_gather_into_tensor(parts, padded, group=self.process_group)
            ranks = torch.cat([parts[r * cap:r * cap + v] for r, v in enumerate(sizes)])
        return ranks

    def _compute_from(self, ranks: torch.Tensor) -> torch.Tensor:
        return ranks

    def compute(self) -> torch.Tensor:
        return self._compute_from(self._all_ranks())

    def compute_local(self) -> torch.Tensor:
        """``compute()`` without the cross-rank "cat": for state that is already global on every rank (the ranks
        :func:`fitclip_b200.retrieval.retrieval_ranks` returns)."""
        if not self.ranks:
            raise RuntimeError("compute_local() called before update()")
        return self._compute_from(torch.cat(self.ranks))

    def reset(self) -> None:
        self.ranks = []
        self.num_candidates = 0

    def clone(self) -> "Rank":
        return type(self)(**self._clone_kwargs())

    def _clone_kwargs(self) -> dict:
        return {"process_group": self.process_group}


class MedianRank(Rank):
    """``ranks.median() + 1`` -- torch's LOWER median (``metrics.py:33-36``), int64."""

    def _compute_from(self, ranks: torch.Tensor) -> torch.Tensor:
        return ops.metrics_from_ranks(ranks.contiguous(), max(self.num_candidates, int(ranks.numel())))[1]


class MeanRank(Rank):
    """``mean(ranks) + 1`` (``metrics.py:27-30``; the reference calls ``.mean()`` on int64, which torch rejects, so
    this is the evident intent in fp32)."""

    def _compute_from(self, ranks: torch.Tensor) -> torch.Tensor:
        return ops.metrics_from_ranks(ranks.contiguous(), max(self.num_candidates, int(ranks.numel())))[2]


class Recall(Rank):
    """Multiclass micro top-k recall == accuracy: ``mean(rank < k)`` as an fp32 fraction in [0, 1]."""
    higher_is_better = True

    def __init__(self, top_k: Optional[int] = None, process_group: Optional[dist.ProcessGroup] = None) -> None:
        super().__init__(process_group)
        self.top_k = top_k or 1

    def update(self, predictions: torch.Tensor, target: torch.Tensor) -> None:
        if self.top_k >= predictions.shape[1] and self.top_k > 1:
            # torchmetrics 0.9 raises too (hence num_sanity_val_steps in config/trainer.yaml:38)
            raise ValueError(f"top_k={self.top_k} must be smaller than the number of candidates "
                             f"({predictions.shape[1]})")
        super().update(predictions, target)

    def _compute_from(self, ranks: torch.Tensor) -> torch.Tensor:
        if self.top_k in (1, 5, 10):
            recall = ops.metrics_from_ranks(ranks.contiguous(), max(self.num_candidates, int(ranks.numel())))[0]
            return recall[{1: 0, 5: 1, 10: 2}[self.top_k]]
        return (ranks < self.top_k).sum().to(torch.float32) / ranks.numel()

    def _clone_kwargs(self) -> dict:
        return {"top_k": self.top_k, "process_group": self.process_group}


Accuracy = Recall
