"""Random initialisation of an OpenAI-layout CLIP state dict -- what ``clip.model.CLIP(**kwargs)`` produces when the
reference instantiates ``config/encoder/clip_from_scratch_vit_b_16.yaml`` (no pretrained weights offline).
Text tower: normal inits of ``initialize_parameters`` (twin ``aligner/encoder/slip.py:438-452``); vision tower blocks:
PyTorch defaults (xavier-uniform ``in_proj_weight``, zero attention biases, kaiming-uniform Linear/Conv), scaled
``randn`` class / positional embeddings and projection."""
from __future__ import annotations

import math
from typing import Dict

import torch


def _linear(out_f: int, in_f: int, g: torch.Generator):
    bound = 1.0 / math.sqrt(in_f)  # kaiming_uniform(a=sqrt(5)) on the weight, uniform(+-1/sqrt(fan_in)) on the bias
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
    return w, b


def _block(sd: Dict[str, torch.Tensor], prefix: str, width: int, g: torch.Generator, text_std=None) -> None:
    if text_std is None:  # vision: nn.MultiheadAttention defaults
        bound = math.sqrt(6.0 / (width + 3 * width))  # xavier_uniform on the packed (3w, w) matrix
        sd[prefix + "attn.in_proj_weight"] = (torch.rand(3 * width, width, generator=g) * 2 - 1) * bound
        sd[prefix + "attn.out_proj.weight"] = _linear(width, width, g)[0]
        fc_w, fc_b = _linear(4 * width, width, g)
        pj_w, pj_b = _linear(width, 4 * width, g)
    else:
        attn_std, proj_std, fc_std = text_std
        sd[prefix + "attn.in_proj_weight"] = torch.randn(3 * width, width, generator=g) * attn_std
        sd[prefix + "attn.out_proj.weight"] = torch.randn(width, width, generator=g) * proj_std
        _, fc_b = _linear(4 * width, width, g)
        _, pj_b = _linear(width, 4 * width, g)
        fc_w = torch.randn(4 * width, width, generator=g) * fc_std
        pj_w = torch.randn(width, 4 * width, generator=g) * proj_std
    sd[prefix + "attn.in_proj_bias"] = torch.zeros(3 * width)
    sd[prefix + "attn.out_proj.bias"] = torch.zeros(width)
    for ln in ("ln_1", "ln_2"):
        sd[prefix + ln + ".weight"], sd[prefix + ln + ".bias"] = torch.ones(width), torch.zeros(width)
    sd[prefix + "mlp.c_fc.weight"], sd[prefix + "mlp.c_fc.bias"] = fc_w, fc_b
    sd[prefix + "mlp.c_proj.weight"], sd[prefix + "mlp.c_proj.bias"] = pj_w, pj_b


def init_clip_state_dict(seed: int = 0, embed_dim: int = 512, image_resolution: int = 224, vision_layers: int = 12,
                         vision_width: int = 768, vision_patch_size: int = 16, context_length: int = 77,
                         vocab_size: int = 49408, transformer_width: int = 512, transformer_heads: int = 8,
                         transformer_layers: int = 12) -> Dict[str, torch.Tensor]:
    assert transformer_heads * 64 == transformer_width, "the native kernels use head dim 64"
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    w, scale = vision_width, vision_width ** -0.5
    tokens = (image_resolution // vision_patch_size) ** 2 + 1
    fan_in = 3 * vision_patch_size ** 2
    sd["visual.class_embedding"] = scale * torch.randn(w, generator=g)
    sd["visual.positional_embedding"] = scale * torch.randn(tokens, w, generator=g)
    sd["visual.proj"] = scale * torch.randn(w, embed_dim, generator=g)
    sd["visual.conv1.weight"] = (torch.rand(w, 3, vision_patch_size, vision_patch_size, generator=g) * 2 - 1) / math.sqrt(fan_in)
    for ln in ("ln_pre", "ln_post"):
        sd[f"visual.{ln}.weight"], sd[f"visual.{ln}.bias"] = torch.ones(w), torch.zeros(w)
    for i in range(vision_layers):
        _block(sd, f"visual.transformer.resblocks.{i}.", w, g)
    t = transformer_width
    stds = (t ** -0.5, (t ** -0.5) * ((2 * transformer_layers) ** -0.5), (2 * t) ** -0.5)
    for i in range(transformer_layers):
        _block(sd, f"transformer.resblocks.{i}.", t, g, text_std=stds)
    sd["token_embedding.weight"] = torch.randn(vocab_size, t, generator=g) * 0.02
    sd["positional_embedding"] = torch.randn(context_length, t, generator=g) * 0.01
    sd["ln_final.weight"], sd["ln_final.bias"] = torch.ones(t), torch.zeros(t)
    sd["text_projection"] = torch.randn(t, embed_dim, generator=g) * t ** -0.5
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


def init_slip_state_dict(seed: int = 0, embed_dim: int = 512, image_resolution: int = 224, vision_layers: int = 12,
                         vision_width: int = 768, vision_patch_size: int = 16, context_length: int = 77,
                         vocab_size: int = 49408, transformer_width: int = 512, transformer_heads: int = 8,
                         transformer_layers: int = 12) -> Dict[str, torch.Tensor]:
    """SLIP-layout state dict as ``CLIP_VITB16()`` leaves it (``aligner/encoder/slip.py:595-600``): timm's
    ``VisionTransformer`` init for the image tower (``trunc_normal_(std=0.02)`` -- cut at +-2 absolute, i.e. a plain normal in
    practice -- for ``pos_embed`` and every Linear weight,
    zero biases, unit LayerNorms, ``cls_token`` ~ N(0, 1e-6); the patch convolution keeps PyTorch's default), the normal
    inits of ``slip.py:438-452`` for the text tower and the two projections."""
    assert transformer_heads * 64 == transformer_width and vision_width % 64 == 0, "widths are multiples of 64"
    g = torch.Generator().manual_seed(seed)

    def trunc(*shape):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=0.02, generator=g)

    sd: Dict[str, torch.Tensor] = {}
    w, t = vision_width, transformer_width
    tokens = (image_resolution // vision_patch_size) ** 2 + 1
    bound = 1.0 / math.sqrt(3 * vision_patch_size ** 2)
    sd["positional_embedding"] = torch.randn(context_length, t, generator=g) * 0.01
    sd["image_projection"] = torch.randn(w, embed_dim, generator=g) * w ** -0.5
    sd["text_projection"] = torch.randn(t, embed_dim, generator=g) * t ** -0.5
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    sd["visual.cls_token"] = torch.randn(1, 1, w, generator=g) * 1e-6
    sd["visual.pos_embed"] = trunc(1, tokens, w)
    sd["visual.patch_embed.proj.weight"] = (torch.rand(w, 3, vision_patch_size, vision_patch_size, generator=g) * 2 - 1) * bound
    sd["visual.patch_embed.proj.bias"] = (torch.rand(w, generator=g) * 2 - 1) * bound
    for i in range(vision_layers):
        p = f"visual.blocks.{i}."
        for name, (o, k) in (("attn.qkv", (3 * w, w)), ("attn.proj", (w, w)), ("mlp.fc1", (4 * w, w)), ("mlp.fc2", (w, 4 * w))):
            sd[p + name + ".weight"], sd[p + name + ".bias"] = trunc(o, k), torch.zeros(o)
        for ln in ("norm1", "norm2"):
            sd[p + ln + ".weight"], sd[p + ln + ".bias"] = torch.ones(w), torch.zeros(w)
    sd["visual.norm.weight"], sd["visual.norm.bias"] = torch.ones(w), torch.zeros(w)
    stds = (t ** -0.5, (t ** -0.5) * ((2 * transformer_layers) ** -0.5), (2 * t) ** -0.5)
    for i in range(transformer_layers):
        _block(sd, f"transformer.resblocks.{i}.", t, g, text_std=stds)
    sd["token_embedding.weight"] = torch.randn(vocab_size, t, generator=g) * 0.02
    sd["ln_final.weight"], sd["ln_final.bias"] = torch.ones(t), torch.zeros(t)
    return sd
