"""Pre-processing hooks returned by the encoder (``aligner/encoder/clip_video_text_encoder.py:113-133``,
``aligner/transforms.py:13-17,56-61``).  They run on CPU inside DataLoader workers, so they are plain picklable
torchvision pipelines; nothing here touches CUDA."""
from __future__ import annotations

import random
from typing import Sequence

import torch
from torch import nn
from torchvision import transforms as T
from torchvision.transforms import InterpolationMode, RandomResizedCrop, functional as F


class ConvertBHWCtoBCHW(nn.Module):
    """(B, H, W, C) -> (B, C, H, W)  (``aligner/transforms.py:13-17``)."""

    def forward(self, v: torch.Tensor) -> torch.Tensor:
        return v.permute(0, 3, 1, 2)


class RandomResizedCropWithRandomInterpolation(RandomResizedCrop):
    """``aligner/transforms.py:56-61``: bilinear or bicubic chosen at random per call."""

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        i, j, h, w = self.get_params(img, self.scale, self.ratio)
        interpolation = random.choice([InterpolationMode.BILINEAR, InterpolationMode.BICUBIC])
        return F.resized_crop(img, i, j, h, w, self.size, interpolation)


def eval_transform(size: int, dtype: torch.dtype, mean: Sequence[float], std: Sequence[float]) -> T.Compose:
    return T.Compose([
        ConvertBHWCtoBCHW(),
        T.ConvertImageDtype(dtype),
        T.Resize(size, interpolation=InterpolationMode.BICUBIC),
        T.CenterCrop(size),
        T.Normalize(mean=mean, std=std),
    ])


def train_transform(size: int, dtype: torch.dtype, mean: Sequence[float], std: Sequence[float]) -> T.Compose:
    return T.Compose([
        ConvertBHWCtoBCHW(),
        T.ConvertImageDtype(dtype),
        RandomResizedCropWithRandomInterpolation(size, scale=(0.5, 1.0)),
        T.RandomHorizontalFlip(),
        T.Normalize(mean=mean, std=std),
    ])
