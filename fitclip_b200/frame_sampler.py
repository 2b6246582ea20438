"""Frame samplers handed to the data layer by the encoder hooks (``aligner/data/frame_sampler.py:12-41``).
Picklable, CPU-only: they run inside DataLoader worker processes."""
from __future__ import annotations

from typing import List

import torch


def _ticks(start_frame: int, end_frame: int, max_frames: int) -> torch.Tensor:
    num_frames = min(max_frames, end_frame - start_frame + 1)
    return torch.linspace(start=start_frame, end=end_frame, steps=num_frames + 1, dtype=torch.int)


class UniformFrameSampler:
    """Midpoint of each of ``max_frames`` equal intervals (``frame_sampler.py:32-41``)."""

    def __init__(self, max_frames: int) -> None:
        self.max_frames = max_frames

    def __call__(self, start_frame: int, end_frame: int, fps: float) -> List[torch.Tensor]:
        t = _ticks(start_frame, end_frame, self.max_frames)
        return [torch.round((a + b) / 2).to(torch.int) for a, b in zip(t[:-1], t[1:])]


class RandomFromUniformIntervalsFrameSampler:
    """A uniformly random frame from each interval (``frame_sampler.py:20-29``)."""

    def __init__(self, max_frames: int) -> None:
        self.max_frames = max_frames

    def __call__(self, start_frame: int, end_frame: int, fps: float) -> List[torch.Tensor]:
        t = _ticks(start_frame, end_frame, self.max_frames)
        return [torch.randint(int(a), int(b) + 1, size=()) for a, b in zip(t[:-1], t[1:])]
