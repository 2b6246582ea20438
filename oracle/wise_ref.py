"""CPU restatement of WiSE weight-space ensembling (``aligner/wise.py:10-23``).
Test infrastructure -- see ``oracle/__init__.py``."""
from __future__ import annotations

import copy
from typing import Mapping, TypeVar

import torch
from torch import nn

T = TypeVar("T", bound=nn.Module)


def ref_wise_state_dict(model1: nn.Module, model2: nn.Module, weight_for_2: float = 0.5) -> Mapping[str, torch.Tensor]:
    # wise.py:10-16 -- python-float scalars times fp32 tensors: two rounded products and a rounded add (no FMA).
    sd1 = dict(model1.named_parameters())
    sd2 = dict(model2.named_parameters())
    assert set(sd1) == set(sd2)
    return {k: (1 - weight_for_2) * sd1[k] + weight_for_2 * sd2[k] for k in sd1}


def ref_wise(model1: T, model2: T, weight_for_2: float = 0.5, copy_model1: bool = True) -> T:
    # wise.py:19-23
    assert type(model1) is type(model2)
    model = copy.deepcopy(model1 if copy_model1 else model2)
    with torch.no_grad():
        model.load_state_dict(ref_wise_state_dict(model1, model2, weight_for_2=weight_for_2))
    return model
