"""Scoring forward of the teacher-student step (``aligner/teacher_student.py:93-96,142-173``): student and frozen-teacher
encoders on the same batch, two scaled ``B x B`` score matrices, NCE on the labelled split and
``KL(softmax(teacher) || softmax(student)) * exp(ts_scale)^2`` on the unlabelled one.  Forward only (SURVEY.md section
8a, a16): the validation side of the module.  The training step -- the same scoring plus backward and AdamW -- is
:class:`fitclip_b200.training.TeacherStudentTrainingModule` (row f3)."""
from __future__ import annotations

import math
from typing import Any, Mapping, Optional, Tuple

import torch
from torch import nn

from . import ops
from .api import TYPE_OUTPUT, VideoTextEncoder
from .retrieval import all_gather_rows


class TeacherStudentScoringModule(nn.Module):
    def __init__(self, encoder: VideoTextEncoder, teacher: VideoTextEncoder, init_temperature: float = 0.05,
                 labeled_dataset_name: str = "labeled", group=None, similarity_terms: int = 3) -> None:
        super().__init__()
        self.encoder = encoder
        self.teacher = teacher
        for p in self.teacher.parameters():  # teacher_student.py:75-76
            p.requires_grad = False
        self.logit_scale = nn.Parameter(torch.tensor([-math.log(init_temperature)]), requires_grad=False)
        self.teacher_student_logit_scale = nn.Parameter(self.logit_scale.clone(), requires_grad=False)  # :70-71
        self.labeled_dataset_name = labeled_dataset_name
        self.group = group
        self.similarity_terms = similarity_terms

    def _step(self, batch: Mapping[str, Any], _batch_idx: int = 0) -> Tuple[TYPE_OUTPUT, TYPE_OUTPUT]:
        # teacher_student.py:93-96
        return (self.encoder(video=batch["video_student"], text=batch["text_student"]),
                self.teacher(video=batch["video_teacher"], text=batch["text_teacher"]))

    def _scores(self, video: torch.Tensor, text: torch.Tensor, scale: float) -> torch.Tensor:
        # `scale * V @ T.T` == (scale * V) @ T.T, rows = videos
        return ops.Similarity(video.contiguous(), text.contiguous(), self.similarity_terms).scores(alpha=scale)

    def _dataset_step_end(self, output: Tuple[TYPE_OUTPUT, TYPE_OUTPUT], split: str = "val",
                          dataset_name: Optional[str] = None) -> torch.Tensor:
        # teacher_student.py:142-173: gather across ranks, then the per-dataset loss
        (video, text), (teacher_video, teacher_text) = output
        video, text, teacher_video, teacher_text = (all_gather_rows(t.contiguous(), self.group)[0]
                                                    for t in (video, text, teacher_video, teacher_text))
        scores = self._scores(video, text, float(self.logit_scale.exp()))
        if dataset_name == self.labeled_dataset_name:
            return ops.nce_loss(scores)
        ts_scale = float(self.teacher_student_logit_scale.exp())
        teacher_scores = self._scores(teacher_video, teacher_text, ts_scale)
        return ops.teacher_student_nce_loss(scores, teacher_scores) * ts_scale ** 2
