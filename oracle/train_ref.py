"""CPU restatement of the teacher-student TRAINING step (autograd + torch.optim.AdamW). Test infrastructure -- see
``oracle/__init__.py``; only ``tests/`` and ``tools/train_step.py``'s checker leg import it.

* ``ref_training_loss``  ``TeacherStudentLightningModule._step`` / ``training_step`` / ``_dataset_step_end`` /
  ``training_step_end`` (``aligner/teacher_student.py:93-96,99-140,142-173,176-183``): student and frozen teacher on
  the same batch; per dataset section ``scores = exp(logit_scale) * V @ T.T``; ``nce_loss`` on the labelled section,
  ``TeacherStudentNCELoss("batchmean")(scores, teacher_scores) * exp(ts_scale)^2`` on the unlabelled one; the sum of
  the section losses weighted by ``dataset_loss_share``.
* ``ref_training_step``  the above + ``loss.backward()`` + ``torch.optim.AdamW(lr=3e-6)`` (``config/trainer.yaml:22-24``,
  ``aligner/cli.py:126-134``; PyTorch defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence, Tuple

import torch

from .loss_ref import ref_nce_loss, ref_teacher_student_nce_loss


def ref_training_loss(student, teacher, batch: Mapping, sections: Sequence[Tuple[str, int, int]],
                      init_temperature: float = 0.05, labeled_dataset_name: str = "labeled",
                      shares: Optional[Mapping[str, float]] = None,
                      prompts: Optional[Tuple[Mapping[str, torch.Tensor], Mapping[str, torch.Tensor]]] = None) -> torch.Tensor:
    """``prompts``: (student-tokenised, teacher-tokenised) prompt lists that replace the texts of the unlabelled section
    (``teacher_student.py:104-120``; same token width as the batch here)."""
    scale = math.exp(-math.log(init_temperature))  # video_text_module.py:32; ts scale is a clone (teacher_student.py:70)
    text_student, text_teacher = batch["text_student"], batch["text_teacher"]
    text_sections = [(lo, hi) for _, lo, hi in sections]
    if prompts is not None:
        idx = next(i for i, (n, _, _) in enumerate(sections) if n != labeled_dataset_name)
        _, lo, hi = sections[idx]
        text_student = {k: torch.cat((v[:lo], prompts[0][k].to(v.dtype), v[hi:])) for k, v in text_student.items()}
        text_teacher = {k: torch.cat((v[:lo], prompts[1][k].to(v.dtype), v[hi:])) for k, v in text_teacher.items()}
        p = len(prompts[0]["input_ids"])
        text_sections = [(a, b) if i < idx else (lo, lo + p) if i == idx else (a + p - (hi - lo), b + p - (hi - lo))
                         for i, (a, b) in enumerate(text_sections)]
    v, t = student(batch["video_student"], text_student)
    with torch.no_grad():
        tv, tt = teacher(batch["video_teacher"], text_teacher)
    names = {n for n, _, _ in sections}
    shares = shares or {n: 1 / max(len(names), 2) for n in names}
    total = 0.0
    for (name, lo, hi), (tlo, thi) in zip(sections, text_sections):
        scores = scale * v[lo:hi] @ t[tlo:thi].T
        if name == labeled_dataset_name:
            loss = ref_nce_loss(scores)
        else:
            teacher_scores = scale * tv[lo:hi] @ tt[tlo:thi].T
            loss = ref_teacher_student_nce_loss(scores, teacher_scores, reduction="batchmean") * scale ** 2
        total = total + loss * shares[name]
    return total


def ref_training_step(student, teacher, batch, sections, optimizer: torch.optim.Optimizer,
                      **kwargs) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """-> (loss, {parameter name: gradient}) after one optimizer step on ``student``."""
    optimizer.zero_grad(set_to_none=True)
    loss = ref_training_loss(student, teacher, batch, sections, **kwargs)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in student.model.named_parameters() if p.grad is not None}
    optimizer.step()
    return loss.detach(), grads
