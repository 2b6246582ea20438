"""The reference's encoder plugin interface, restated so the package imports without ``overrides``/Lightning:
``VideoEncoder`` (``aligner/encoder/video_encoder.py:14-52``) and ``VideoTextEncoder``
(``aligner/encoder/video_text_encoder.py:15-31``).  Same method names, argument meaning and error behaviour
(un-implemented hooks raise ``NotImplementedError``)."""
from __future__ import annotations

from abc import abstractmethod
from typing import Callable, Iterable, Iterator, Mapping, Optional, Sequence, Tuple

import torch
from torch import nn

TYPE_VIDEO_INPUT = torch.Tensor
TYPE_TRANSFORM = Callable[[torch.Tensor], torch.Tensor]
TYPE_TEXT_INPUT = Mapping[str, torch.Tensor]
TYPE_OUTPUT = Tuple[torch.Tensor, torch.Tensor]
TYPE_TOKENIZER = Callable[[Iterable[str]], Mapping[str, torch.Tensor]]
FrameSampler = Callable[[int, int, float], Sequence[int]]  # aligner/data/frame_sampler.py:12-17


class VideoEncoder(nn.Module):
    @abstractmethod
    def encode_video(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        return self.encode_video(video)

    @abstractmethod
    def get_train_frame_sampler(self) -> FrameSampler:
        raise NotImplementedError

    @abstractmethod
    def get_eval_frame_sampler(self) -> FrameSampler:
        raise NotImplementedError

    @abstractmethod
    def get_train_transform(self, dtype: torch.dtype) -> TYPE_TRANSFORM:
        raise NotImplementedError

    @abstractmethod
    def get_eval_transform(self, dtype: torch.dtype) -> TYPE_TRANSFORM:
        raise NotImplementedError

    @property
    def should_pad_batch(self) -> bool:
        raise NotImplementedError

    @abstractmethod
    def to_bchw(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def denormalize_video_tensor(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        """Converts a transformed video tensor into an unsigned 8-bit integer tensor in the range 0-255."""
        raise NotImplementedError


def float_standard_denormalize(video: TYPE_VIDEO_INPUT, mean: Optional[Tuple[float, float, float]] = None,
                               std: Optional[Tuple[float, float, float]] = None) -> torch.Tensor:
    # aligner/encoder/video_encoder.py:55-63 (in place on `video`, like the reference)
    if std is not None:
        video *= torch.tensor(std, device=video.device, dtype=video.dtype).view(-1, 1, 1)
    if mean is not None:
        video += torch.tensor(mean, device=video.device, dtype=video.dtype).view(-1, 1, 1)
    return (video * 255).to(torch.uint8)


class VideoTextEncoder(VideoEncoder):
    @abstractmethod
    def encode_text(self, text: TYPE_TEXT_INPUT) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, video: TYPE_VIDEO_INPUT, text: TYPE_TEXT_INPUT) -> TYPE_OUTPUT:  # noqa
        return self.encode_video(video), self.encode_text(text)

    @abstractmethod
    def get_tokenizer(self) -> TYPE_TOKENIZER:
        raise NotImplementedError

    @abstractmethod
    def decode_text(self, text: TYPE_TEXT_INPUT) -> Iterator[str]:
        """Decodes a batch of texts."""
        raise NotImplementedError
