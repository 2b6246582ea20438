"""CPU restatement of the reference's evaluation pre-processing (``ClipVideoTextEncoder.get_eval_transform``,
``aligner/encoder/clip_video_text_encoder.py:124-133`` + ``aligner/transforms.py:13-17``) in plain torch.
Test infrastructure -- see ``oracle/__init__.py``.

The reference pins torchvision 0.12 (``environment.yml``), whose ``Resize`` does NOT antialias tensor inputs
(``antialias=None`` -> False in ``functional_tensor.resize``) and sizes the output as
``new_short, new_long = size, int(size * long / short)``; ``CenterCrop`` offsets are ``int(round((h - crop) / 2.0))``
(Python round-half-even).  Pinned in ``tests/test_oracle_preprocess.py`` against the torchvision in this image with
``antialias=False``."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F


def ref_resized_size(h: int, w: int, size: int):
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def ref_eval_transform(video: torch.Tensor, size: int, mean: Sequence[float], std: Sequence[float],
                       dtype: torch.dtype = torch.float32, interpolation: str = "bicubic") -> torch.Tensor:
    """uint8 ``(T, H, W, 3)`` -> ``dtype`` ``(T, 3, size, size)``.  ``interpolation="bilinear"``: the SLIP wrapper's
    transform, which keeps ``Resize``'s default (``slip_video_text_encoder.py:78-87``)."""
    assert video.dtype == torch.uint8 and video.shape[-1] == 3
    x = video.permute(0, 3, 1, 2)                       # ConvertBHWCtoBCHW (aligner/transforms.py:13-17)
    x = x.to(dtype) / 255                               # ConvertImageDtype(uint8 -> float)
    h, w = x.shape[-2:]
    new_h, new_w = ref_resized_size(h, w, size)
    if (h, w) != (new_h, new_w):                        # Resize(size, BICUBIC): shorter side -> size, no antialias
        x = F.interpolate(x, size=(new_h, new_w), mode=interpolation, align_corners=False, antialias=False)
    top = int(round((new_h - size) / 2.0))              # CenterCrop(size)
    left = int(round((new_w - size) / 2.0))
    x = x[..., top:top + size, left:left + size]
    m = torch.tensor(mean, dtype=dtype).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=dtype).view(1, 3, 1, 1)
    return (x - m) / s                                  # Normalize(mean, std)
