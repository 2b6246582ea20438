"""Literal CPU restatement of ``ClipVideoTextEncoder.encode_video/encode_text``
(``aligner/encoder/clip_video_text_encoder.py:80-94``) and of the similarity / classification call sites.
Test infrastructure -- see ``oracle/__init__.py``."""
from __future__ import annotations

from typing import Mapping, Tuple

import torch
from torch import nn


class RefClipVideoTextEncoder(nn.Module):
    def __init__(self, model: nn.Module, num_frames: int = 4) -> None:
        super().__init__()
        self.model = model
        self.num_frames = num_frames
        if hasattr(self.model, "logit_scale"):  # clip_video_text_encoder.py:75-77
            delattr(self.model, "logit_scale")

    def encode_video(self, video: torch.Tensor) -> torch.Tensor:
        # clip_video_text_encoder.py:80-89: flatten frames, encode, L2-normalise per FRAME, mean over frames
        # (the mean is NOT re-normalised).
        batch_size = video.shape[0]
        images = video.view(-1, *video.shape[2:])
        encoded = self.model.encode_image(images)
        encoded = encoded / encoded.norm(dim=-1, keepdim=True)
        return encoded.view(batch_size, -1, *encoded.shape[1:]).mean(dim=1)

    def encode_text(self, text: Mapping[str, torch.Tensor]) -> torch.Tensor:
        # clip_video_text_encoder.py:92-94
        encoded = self.model.encode_text(text["input_ids"])
        return encoded / encoded.norm(dim=-1, keepdim=True)

    def forward(self, video: torch.Tensor, text: Mapping[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        # aligner/encoder/video_text_encoder.py:20-22
        return self.encode_video(video), self.encode_text(text)


def ref_batch_scores(encoded_video: torch.Tensor, encoded_text: torch.Tensor, logit_scale: float) -> torch.Tensor:
    """Per-batch scaled scores: ``logit_scale * V @ T.T`` parses as ``(logit_scale * V) @ T.T``
    (``aligner/text_video_retrieval.py:49-50``, ``aligner/video_text_module.py:62-63``)."""
    return logit_scale * encoded_video @ encoded_text.T


def ref_retrieval_scores(encoded_texts: torch.Tensor, encoded_videos: torch.Tensor) -> torch.Tensor:
    """Epoch-level UNSCALED similarity, rows = texts, columns = videos (``aligner/text_video_retrieval.py:74``)."""
    return encoded_texts @ encoded_videos.T


def ref_class_embeddings(encoder: RefClipVideoTextEncoder, input_ids: torch.Tensor, template_count: int,
                         batch_size: int = 32) -> torch.Tensor:
    """``VideoTextClassificationLightningModule._on_start`` (``aligner/video_text_classification.py:69-96``):
    encode label x template prompts in batches of 32, mean over templates, no re-normalisation."""
    chunks = [encoder.encode_text({"input_ids": input_ids[i:i + batch_size]})
              for i in range(0, input_ids.shape[0], batch_size)]
    encoded = torch.cat(chunks)
    return encoded.reshape(-1, template_count, encoded.shape[1]).mean(dim=1)
