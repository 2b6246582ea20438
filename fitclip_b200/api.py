"""Encoder plugin surface of the reference, so that code written against ``aligner.encoder.video_encoder.VideoEncoder``
(``aligner/encoder/video_encoder.py:14-52``) / ``aligner.encoder.video_text_encoder.VideoTextEncoder``
(``aligner/encoder/video_text_encoder.py:15-31``) finds the same names here without ``overrides`` or Lightning.

The contract is a list of hook names (below): a hook a subclass does not provide raises ``NotImplementedError`` when
called -- the reference's error behaviour -- and ``forward`` is ``encode_video`` (``+ encode_text`` for the video-text
variant).  The hooks are installed from that list rather than spelled out one by one."""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Mapping, Optional, Sequence, Tuple

import torch
from torch import nn

TYPE_VIDEO_INPUT = torch.Tensor                                             # (B, T, 3, H, W)
TYPE_TEXT_INPUT = Mapping[str, torch.Tensor]                                # {"input_ids": (C, context_length)}
TYPE_OUTPUT = Tuple[torch.Tensor, torch.Tensor]                             # (video embeddings, text embeddings)
TYPE_TRANSFORM = Callable[[torch.Tensor], torch.Tensor]                     # uint8 (T, H, W, 3) -> float (T, 3, R, R)
TYPE_TOKENIZER = Callable[[Iterable[str]], Mapping[str, torch.Tensor]]
FrameSampler = Callable[[int, int, float], Sequence[int]]                   # (start_frame, end_frame, fps) -> indices

# hook name -> what it returns (data modules pull these from the encoder: aligner/data/video_data_module.py:40-55,74-78)
VIDEO_HOOKS = {
    "encode_video": "(B, T, 3, H, W) frames -> (B, E) embeddings",
    "get_train_frame_sampler": "FrameSampler used for training clips",
    "get_eval_frame_sampler": "FrameSampler used for evaluation clips",
    "get_train_transform": "dtype -> TYPE_TRANSFORM with augmentation",
    "get_eval_transform": "dtype -> deterministic TYPE_TRANSFORM",
    "to_bchw": "the transformed video laid out (B, C, H, W)",
    "denormalize_video_tensor": "a transformed video back to uint8 pixels in 0..255",
}
TEXT_HOOKS = {
    "encode_text": "{'input_ids': (C, L)} -> (C, E) embeddings",
    "get_tokenizer": "TYPE_TOKENIZER",
    "decode_text": "token ids back to an iterator of strings",
}


def _missing_hook(name: str, doc: str):
    def hook(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} does not implement {name}() [{doc}]")

    hook.__name__ = hook.__qualname__ = name
    hook.__doc__ = doc
    hook.__isabstractmethod__ = True
    return hook


def _install(cls, hooks: Mapping[str, str]) -> None:
    for name, doc in hooks.items():
        setattr(cls, name, _missing_hook(name, doc))


class VideoEncoder(nn.Module):
    """Video-only encoder plugin: subclasses provide the hooks in ``VIDEO_HOOKS``."""

    def forward(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        return self.encode_video(video)

    @property
    def should_pad_batch(self) -> bool:  # whether collate pads clips of different lengths (video_dataset.py:102-112)
        raise NotImplementedError(f"{type(self).__name__} does not say whether batches are padded")


class VideoTextEncoder(VideoEncoder):
    """Video + text encoder plugin: additionally the hooks in ``TEXT_HOOKS``; ``forward`` returns both embeddings."""

    def forward(self, video: TYPE_VIDEO_INPUT, text: TYPE_TEXT_INPUT) -> TYPE_OUTPUT:  # noqa: signature differs by design
        return self.encode_video(video), self.encode_text(text)


_install(VideoEncoder, VIDEO_HOOKS)
_install(VideoTextEncoder, TEXT_HOOKS)


def float_standard_denormalize(video: TYPE_VIDEO_INPUT, mean: Optional[Sequence[float]] = None,
                               std: Optional[Sequence[float]] = None) -> torch.Tensor:
    """Inverse of ``Normalize(mean, std)`` followed by the 0..255 quantisation.  Works IN PLACE on ``video`` like the
    reference helper of the same name (``aligner/encoder/video_encoder.py:55-63``), channel axis third from the end."""
    def per_channel(values: Sequence[float]) -> torch.Tensor:
        return torch.as_tensor(values, dtype=video.dtype, device=video.device).reshape(-1, 1, 1)

    if std is not None:
        video.mul_(per_channel(std))
    if mean is not None:
        video.add_(per_channel(mean))
    return video.mul(255).to(torch.uint8)
