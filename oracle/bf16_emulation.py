"""CPU emulation of a bf16 residual stream on top of the fp32 oracle (test infrastructure -- see ``oracle/__init__.py``).

The CUDA path keeps the residual stream, the LayerNorm-ed GEMM inputs and the block outputs in bf16 (fp32 accumulation
and statistics).  ``bf16_stream_model`` rounds the oracle's activations to bf16 at those points and
nowhere else, so ``error(emulation vs fp32 oracle)`` is the error bf16 storage alone explains; the GPU parity tests
bound ``error(CUDA vs fp32 oracle)`` by a small multiple of it -- a LayerNorm-folding or bias bug shows up as an error
the storage format does not explain, whatever the weights' conditioning."""
from __future__ import annotations

import copy

import torch
from torch import nn


def _rb(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def bf16_stream_model(model: nn.Module) -> nn.Module:
    """Deep copy of an oracle ``CLIP`` whose residual stream / block inputs / block outputs are rounded to bf16."""
    m = copy.deepcopy(model)
    for tower in (m.visual.transformer, m.transformer):
        for block in tower.resblocks:
            def forward(x, block=block):
                x = _rb(x + _rb(block.attention(_rb(block.ln_1(x)))))
                return _rb(x + _rb(block.mlp(_rb(block.ln_2(x)))))
            block.forward = forward
    m.visual.ln_pre.register_forward_pre_hook(lambda mod, args: (_rb(args[0]),))
    m.visual.ln_pre.register_forward_hook(lambda mod, args, out: _rb(out))
    # the text stream starts at token + positional embedding (no ln_pre): stored in bf16 as well
    m.transformer.register_forward_pre_hook(lambda mod, args: (_rb(args[0]),))
    return m
