"""fp32 CPU restatement of ``clip.model.CLIP`` (openai/CLIP pinned at b46f5ac by the reference's
``environment.yml:7``; the package itself is NOT vendored under /root/reference and NOT installed here).

Test infrastructure -- see ``oracle/__init__.py``. What each piece follows:

* ``LayerNorm`` / ``QuickGELU`` / ``ResidualAttentionBlock`` / ``Transformer``: the in-tree twin
  ``aligner/encoder/slip.py:350-396`` (fp32 LayerNorm with cast-back ``:350-356``; ``x * sigmoid(1.702 x)``
  ``:359-361``; pre-LN residual block around ``nn.MultiheadAttention`` ``:364-385``).
* text tower, causal mask, EOT pooling, text-tower init: ``aligner/encoder/slip.py:399-480``
  (mask ``:454-460``, ``encode_text`` ``:468-480``, init ``:438-452``).
* vision tower (``VisionTransformer``), ``build_model`` and the vision init are third-party-only; restated from
  the published algorithm and cross-checked against ``transformers.CLIPModel`` in
  ``tests/test_oracle_clip.py``.
* constructor keywords: ``config/encoder/clip_from_scratch_vit_b_16.yaml:5-16``.

State-dict names are the OpenAI ones (``visual.conv1.weight``, ``transformer.resblocks.N.attn.in_proj_weight``,
``token_embedding.weight``, ``positional_embedding``, ``ln_final.*``, ``text_projection``, ``logit_scale``), which the
reference pins in ``config/trainer/callbacks/clip_freeze_text.yaml:25-29``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Mapping, Optional

import numpy as np
import torch
from torch import nn

SOT_TOKEN = 49406
EOT_TOKEN = 49407


class LayerNorm(nn.LayerNorm):
    # slip.py:350-356 -- statistics in fp32 whatever the input type, then cast back.
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return super().forward(x.to(torch.float32)).to(x.dtype)


class QuickGELU(nn.Module):
    # slip.py:359-361
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    # slip.py:364-385
    def __init__(self, d_model: int, n_head: int, attn_mask: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor) -> torch.Tensor:
        mask = None if self.attn_mask is None else self.attn_mask.to(dtype=x.dtype, device=x.device)
        return self.attn(x, x, x, need_weights=False, attn_mask=mask)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    # slip.py:388-396
    def __init__(self, width: int, layers: int, heads: int, attn_mask: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    """[3P] ViT image tower: stride-16 conv patch embedding, class token, learned positions, ln_pre,
    seq-first transformer, ln_post on the class token, projection."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int,
                 output_dim: int) -> None:
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv1(x)                                    # (F, width, g, g)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # (F, g*g, width)
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1)                       # (F, g*g+1, width)
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)                               # NLD -> LND
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class CLIP(nn.Module):
    """Keyword-compatible with ``clip.model.CLIP`` as instantiated by
    ``config/encoder/clip_from_scratch_vit_b_16.yaml:5-16`` (ViT towers only: ``vision_layers`` is an int)."""

    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int,
                 vision_patch_size: int, context_length: int, vocab_size: int, transformer_width: int,
                 transformer_heads: int, transformer_layers: int) -> None:
        super().__init__()
        self.context_length = context_length
        self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size,
                                        width=vision_width, layers=vision_layers, heads=vision_width // 64,
                                        output_dim=embed_dim)
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self) -> None:
        # slip.py:438-452 (text side); the vision blocks keep the PyTorch defaults [3P].
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self) -> torch.Tensor:
        # slip.py:454-460 -- additive mask, -inf strictly above the diagonal.
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self) -> torch.dtype:
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image.to(self.dtype))

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        # slip.py:468-480
        x = self.token_embedding(text).to(self.dtype)
        x = x + self.positional_embedding.to(self.dtype)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_final(x).to(self.dtype)
        # the EOT token has the largest id in each sequence
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection


def build_model(state_dict: Mapping[str, torch.Tensor]) -> CLIP:
    """[3P] ``clip.model.build_model`` for ViT checkpoints: every dimension is inferred from tensor shapes
    (SURVEY.md Appendix A). Kept in fp32 (the reference forces ``model.float()``,
    ``aligner/encoder/clip_video_text_encoder.py:22-25``)."""
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    vision_patch_size = state_dict["visual.conv1.weight"].shape[-1]
    grid_size = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    embed_dim = state_dict["text_projection"].shape[1]
    context_length = state_dict["positional_embedding"].shape[0]
    vocab_size = state_dict["token_embedding.weight"].shape[0]
    transformer_width = state_dict["ln_final.weight"].shape[0]
    transformer_layers = len({k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")})
    model = CLIP(embed_dim, vision_patch_size * grid_size, vision_layers, vision_width, vision_patch_size,
                 context_length, vocab_size, transformer_width, transformer_width // 64, transformer_layers)
    state_dict = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    model.load_state_dict(state_dict)
    return model.float().eval()


VIT_B_16 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                transformer_layers=12)  # config/encoder/clip_from_scratch_vit_b_16.yaml:7-16


def perturb_trained_like(model: nn.Module, seed: int = 0, stress: bool = False) -> nn.Module:
    """Moves the parameters that every init path leaves at an identity value (LayerNorm gamma = 1 / beta = 0,
    ``in_proj_bias`` = ``out_proj.bias`` = 0) to where a trained checkpoint has them, so that parity tests exercise the
    LayerNorm-folding arithmetic (``W diag(gamma)``, ``b + W beta``, the column-sum term) and the attention bias adds
    with non-trivial values: gamma ~ U(0.2, 3), beta ~ N(0, 0.5), attention biases ~ N(0, 0.1).

    ``stress`` adds what pretrained CLIP checkpoints are known for, at magnitudes a bf16 residual stream still carries
    (calibrated with a CPU emulation that rounds the stream to bf16: cosine >= 0.9995 vs fp32; putting the x10 gammas on
    the SAME channels as the x60 writers compounds to a x600 loop gain per block that not even the reference's fp16 GPU
    path survives): four "massive activation" channels per tower (the ``out_proj`` / ``c_proj`` rows that write them
    x60, their LayerNorm gammas x0.3 as trained models have them), four other channels with gamma x10 on every
    LayerNorm of the stream, and a DC offset on three token rows before the first LayerNorm (+20 on positional-embedding
    rows of the image tower, whose stream passes ``ln_pre`` first; +0.1 = 5 sigma of the token embeddings on the text
    tower, whose bf16 stream starts at the raw 0.02-scale embeddings: a DC of 200 sigma there (+4) cannot be stored in 8
    mantissa bits, the emulation fails it exactly like the kernels do) -- what punishes an ``E[x^2] - mean^2`` variance.  Test infrastructure (``oracle/__init__.py``); in place."""
    g = torch.Generator().manual_seed(1_000_003 + seed)

    def rand(shape, lo, hi):
        return lo + (hi - lo) * torch.rand(shape, generator=g)

    def randn(shape, std):
        return std * torch.randn(shape, generator=g)

    with torch.no_grad():
        for name, p in model.named_parameters():
            parent, _, leaf = name.rpartition(".")
            is_ln = parent.rsplit(".", 1)[-1] in ("ln_1", "ln_2", "ln_pre", "ln_post", "ln_final")
            if is_ln and leaf == "weight":
                p.copy_(rand(p.shape, 0.2, 3.0))
            elif is_ln and leaf == "bias":
                p.copy_(randn(p.shape, 0.5))
            elif name.endswith("attn.in_proj_bias") or name.endswith("attn.out_proj.bias"):
                p.copy_(randn(p.shape, 0.1))
        if stress:
            for visual, width, dc in ((True, model.visual.conv1.weight.shape[0], 20.0),
                                      (False, model.transformer.width, 0.1)):
                perm = torch.randperm(width, generator=g)
                gamma_channels, massive_channels = perm[:4], perm[4:8]
                for name, p in model.named_parameters():
                    if name.startswith("visual.") != visual:
                        continue
                    parent, _, leaf = name.rpartition(".")
                    if parent.rsplit(".", 1)[-1] in ("ln_1", "ln_2", "ln_pre") and leaf == "weight":
                        p[gamma_channels] *= 10.0
                        p[massive_channels] *= 0.3
                    elif name.endswith("attn.out_proj.weight") or name.endswith("mlp.c_proj.weight"):
                        p[massive_channels] *= 60.0
                pos = model.visual.positional_embedding if visual else model.positional_embedding
                rows = torch.randperm(pos.shape[0], generator=g)[:3]
                pos[rows] += dc
    return model


def clip_vit_b_16(seed: int = 0, trained_like: bool = True, stress: bool = False, **overrides) -> CLIP:
    """Random-init ViT-B/16 (the benchmark weights: no network, no checkpoints). ``overrides`` shrink it for tests.
    ``trained_like`` (default) applies :func:`perturb_trained_like` on top of the reference's init so that no LayerNorm
    or attention bias sits at its identity value; ``trained_like=False`` is the bare ``CLIP.__init__`` state
    (``config/encoder/clip_from_scratch_vit_b_16.yaml``)."""
    torch.manual_seed(seed)
    model = CLIP(**{**VIT_B_16, **overrides}).float().eval()
    if trained_like:
        perturb_trained_like(model, seed, stress)
    return model


def tokenize_synthetic(count: int, length: int | tuple[int, int] = 77, seed: int = 4321, context_length: int = 77,
                       vocab_size: int = 49408) -> torch.Tensor:
    """Synthetic stand-in for ``clip.tokenize(texts, truncate=True)`` (the BPE vocabulary file is not on disk):
    ``[SOT] + random ids + [EOT]`` zero-padded to ``context_length``, int32.  ``length`` is the
    total token count including SOT/EOT, or an inclusive (lo, hi) range for ragged captions."""
    g = torch.Generator().manual_seed(seed)
    eot = vocab_size - 1
    sot = vocab_size - 2
    ids = torch.zeros(count, context_length, dtype=torch.int32)
    if isinstance(length, tuple):
        lengths = torch.randint(length[0], length[1] + 1, (count,), generator=g)
    else:
        lengths = torch.full((count,), length)
    body = torch.randint(1, sot, (count, context_length), generator=g, dtype=torch.int32)
    for i in range(count):
        n = int(lengths[i])
        ids[i, :n] = body[i, :n]
        ids[i, 0] = sot
        ids[i, n - 1] = eot
    return ids
