"""WiSE weight-space ensembling at load (``aligner/wise.py:10-23``) on the single-pass lerp kernel."""
from __future__ import annotations

import copy
from typing import Mapping, TypeVar

import torch
from torch import nn

from . import ops

T = TypeVar("T", bound=nn.Module)


def wise_state_dict(model1: T, model2: T, weight_for_2: float = 0.5) -> Mapping[str, torch.Tensor]:
    """``{k: (1 - w) * p1[k] + w * p2[k]}`` over ``named_parameters()`` (``wise.py:10-16``); bit-exact with the
    reference's torch expression (two rounded products + rounded add, no FMA).  The lerp always runs on the GPU;
    parameters held on the CPU (the reference loads models with ``device="cpu"``,
    ``aligner/encoder/clip_video_text_encoder.py:55-56``) are staged through the current CUDA device and returned on
    their own device."""
    sd1 = dict(model1.named_parameters())
    sd2 = dict(model2.named_parameters())
    assert set(sd1) == set(sd2)
    out = {}
    for k, p1 in sd1.items():
        p2 = sd2[k]
        if p1.dtype != torch.float32 or p2.dtype != torch.float32:
            raise TypeError(f"WiSE expects fp32 parameters, {k} is {p1.dtype}/{p2.dtype}")
        dev = p1.device if p1.is_cuda else torch.device("cuda", torch.cuda.current_device())
        out[k] = ops.wise_lerp(p1.detach().to(dev).contiguous(), p2.detach().to(dev).contiguous(),
                               weight_for_2).to(p1.device)
    return out


def wise(model1: T, model2: T, weight_for_2: float = 0.5, copy_model1: bool = True) -> T:
    """``wise.py:19-23``: deep-copy one model and strictly load the interpolated parameters into it."""
    assert type(model1) is type(model2)
    model = copy.deepcopy(model1 if copy_model1 else model2)
    with torch.no_grad():
        model.load_state_dict(wise_state_dict(model1, model2, weight_for_2=weight_for_2))
    return model
