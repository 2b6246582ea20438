"""Pre-processing hooks returned by the encoder (``aligner/encoder/clip_video_text_encoder.py:113-133``,
``aligner/transforms.py:13-17,56-61``).  They run on CPU inside DataLoader workers, so they are plain picklable
torchvision pipelines; nothing here touches CUDA.  (The GPU twin of the eval pipeline is
``fitclip_b200.ops.preprocess_frames`` / ``B200ClipVideoTextEncoder.encode_video_uint8``.)

``antialias=False`` is spelled out: the reference pins torchvision 0.12, which never antialiases tensor inputs, while
current torchvision defaults to antialiasing -- the explicit flag keeps the reference's pixels on any version."""
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
        return F.resized_crop(img, i, j, h, w, self.size, interpolation, antialias=False)


def eval_transform(size: int, dtype: torch.dtype, mean: Sequence[float], std: Sequence[float],
                   interpolation: InterpolationMode = InterpolationMode.BICUBIC) -> T.Compose:
    # bicubic: clip_video_text_encoder.py:124-133; the SLIP wrapper keeps Resize's default, bilinear
    # (slip_video_text_encoder.py:78-87)
    return T.Compose([
        ConvertBHWCtoBCHW(),
        T.ConvertImageDtype(dtype),
        T.Resize(size, interpolation=interpolation, antialias=False),
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
