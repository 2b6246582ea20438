"""CPU restatement of the per-batch losses (forward only). Test infrastructure -- see ``oracle/__init__.py``.

* ``ref_nce_loss``                 ``aligner/loss.py:13-26``
* ``ref_teacher_student_nce_loss`` ``aligner/loss.py:29-39`` (module default reduction is "mean";
  ``kl_div(..., reduction="mean")`` divides by the element count B*B)
"""
from __future__ import annotations

import torch
from torch.nn import functional as F


def _rows(scores: torch.Tensor, reduction: str) -> torch.Tensor:
    loss = -F.log_softmax(scores, dim=-1).diag()
    return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss


def ref_nce_loss(scores: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
    return _rows(scores, reduction) + _rows(scores.T, reduction)


def _ts_rows(scores: torch.Tensor, teacher_scores: torch.Tensor, reduction: str) -> torch.Tensor:
    return F.kl_div(F.log_softmax(scores, dim=-1), F.softmax(teacher_scores, dim=-1), reduction=reduction)


def ref_teacher_student_nce_loss(scores: torch.Tensor, teacher_scores: torch.Tensor,
                                 reduction: str = "mean") -> torch.Tensor:
    return _ts_rows(scores, teacher_scores, reduction) + _ts_rows(scores.T, teacher_scores.T, reduction)
