"""TEST INFRASTRUCTURE -- CPU restatement of the SLIP-layout models of the reference.

The reference builds these as ``aligner/encoder/slip.py:399-480`` (class ``CLIP``: an arbitrary ``vision_model`` +
``image_projection`` + CLIP's own text tower) with ``vision_model = timm.create_model('vit_base_patch16_224',
num_classes=0)`` (``slip.py:595-600``; ``vit_large_patch16_224`` ``:618-623``; ``SLIP_*`` add SSL heads that the
evaluation never calls, ``:530-557``).  ``timm`` is a third-party dependency that is NOT vendored in the reference and
not installed here (``environment.yml:43``, un-pinned ``conda-forge::timm``; the 0.5 / 0.6 series of 2022), so its
``VisionTransformer`` is restated below from its published algorithm:

    x = patch_embed.proj(img)            Conv2d(3, W, 16, stride 16, bias=True) -> (B, W, 14, 14) -> (B, 196, W)
    x = cat(cls_token, x) + pos_embed    cls_token (1, 1, W), pos_embed (1, 197, W); no dropout at eval
    for blk in blocks:                   pre-LN, LayerNorm eps 1e-6, qkv_bias=True, mlp_ratio 4, exact (erf) nn.GELU
        x = x + attn(norm1(x)); x = x + mlp(norm2(x))
    x = norm(x)[:, 0]                    global_pool='token', head = Identity (num_classes=0)

with timm's parameter names, so that a SLIP checkpoint's state dict loads strictly.  It is pinned against an
independent implementation of the same architecture, ``transformers.ViTModel`` (tests/test_oracle_slip.py), and the
wrapper / projection / text tower around it against the reference's own ``slip.CLIP`` and ``SlipVideoTextEncoder`` run in
the build container (tests/golden/make_reference_slip_golden.py -> tests/golden/reference_slip.pt).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .clip_ref import LayerNorm, Transformer

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # slip_video_text_encoder.py:91
IMAGENET_STD = (0.229, 0.224, 0.225)


class TimmAttention(nn.Module):
    """[3P] timm.models.vision_transformer.Attention."""

    def __init__(self, dim: int, num_heads: int) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class TimmMlp(nn.Module):
    """[3P] timm.models.layers.Mlp with the default ``act_layer=nn.GELU`` (exact, erf)."""

    def __init__(self, dim: int, hidden: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(self.act(self.fc1(x)))


class TimmBlock(nn.Module):
    """[3P] timm.models.vision_transformer.Block (no layer scale, no drop path at eval)."""

    def __init__(self, dim: int, num_heads: int) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = TimmAttention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = TimmMlp(dim, 4 * dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, patch: int, width: int) -> None:
        super().__init__()
        self.proj = nn.Conv2d(3, width, kernel_size=patch, stride=patch)  # bias=True, unlike CLIP's conv1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj(x).flatten(2).transpose(1, 2)


class TimmVisionTransformer(nn.Module):
    """[3P] ``timm.create_model('vit_*_patch16_224', num_classes=0)``: returns the normalised class token, (B, W)."""

    def __init__(self, img_size: int = 224, patch_size: int = 16, embed_dim: int = 768, depth: int = 12,
                 num_heads: int = 12) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.patch_embed = _PatchEmbed(patch_size, embed_dim)
        tokens = (img_size // patch_size) ** 2 + 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, tokens, embed_dim))
        self.blocks = nn.Sequential(*[TimmBlock(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        # timm's init: trunc_normal(std .02) for pos_embed and Linear weights, normal(std 1e-6) for cls_token, zero biases
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        return self.norm(self.blocks(x))[:, 0]


class SlipClip(nn.Module):
    """``slip.py:399-480`` (class CLIP): vision_model + image_projection, CLIP's text tower, same parameter names."""

    def __init__(self, embed_dim: int, vision_width: int, vision_model: nn.Module, context_length: int, vocab_size: int,
                 transformer_width: int, transformer_heads: int, transformer_layers: int) -> None:
        super().__init__()
        self.context_length = context_length
        self.vision_width = vision_width
        self.visual = vision_model
        mask = torch.full((context_length, context_length), float("-inf")).triu_(1)  # slip.py:454-460
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads, attn_mask=mask)
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.image_projection = nn.Parameter(torch.empty(vision_width, embed_dim))
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        # slip.py:438-452
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (transformer_width ** -0.5) * ((2 * transformer_layers) ** -0.5)
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=transformer_width ** -0.5)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=(2 * transformer_width) ** -0.5)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.image_projection, std=vision_width ** -0.5)
        nn.init.normal_(self.text_projection, std=transformer_width ** -0.5)

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:  # slip.py:462-466
        return self.visual(image) @ self.image_projection

    def encode_text(self, text: torch.Tensor) -> torch.Tensor:  # slip.py:468-480
        x = self.token_embedding(text) + self.positional_embedding
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        x = self.ln_final(x)
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection


class RefSlipVideoTextEncoder(nn.Module):
    """``SlipVideoTextEncoder.encode_video / encode_text`` (slip_video_text_encoder.py:37-51)."""

    def __init__(self, model: SlipClip, num_frames: int = 4) -> None:
        super().__init__()
        self.model = model
        self.num_frames = num_frames

    def encode_video(self, video: torch.Tensor) -> torch.Tensor:
        batch_size = video.shape[0]
        x = self.model.encode_image(video.view(-1, *video.shape[2:]))
        x = x / x.norm(dim=-1, keepdim=True)
        return x.view(batch_size, -1, *x.shape[1:]).mean(dim=1)

    def encode_text(self, text) -> torch.Tensor:
        x = self.model.encode_text(text["input_ids"])
        return x / x.norm(dim=-1, keepdim=True)


def perturb_timm_trained_like(vit: TimmVisionTransformer, seed: int) -> None:
    """Non-trivial values wherever timm's init leaves identities or zeros (LayerNorm gamma / beta, every bias, the class
    token): the same ranges as ``clip_ref.perturb_trained_like``."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in vit.modules():
            if isinstance(m, nn.LayerNorm):
                m.weight.copy_(torch.empty_like(m.weight).uniform_(0.2, 3.0, generator=g))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.5)
            elif isinstance(m, nn.Linear):
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        vit.patch_embed.proj.bias.copy_(torch.randn(vit.patch_embed.proj.bias.shape, generator=g) * 0.1)
        vit.cls_token.copy_(torch.randn(vit.cls_token.shape, generator=g) * 0.02)


def slip_clip_vit_b_16(seed: int = 0, trained_like: bool = True, **overrides) -> SlipClip:
    """``CLIP_VITB16`` (slip.py:595-600), random init under ``seed``.  ``overrides``: ``img_size``, ``patch_size``,
    ``vision_width``, ``vision_layers``, ``vision_heads`` for the tower and the ``SlipClip`` text-side arguments."""
    from .clip_ref import perturb_trained_like
    cfg = dict(embed_dim=512, vision_width=768, context_length=77, vocab_size=49408, transformer_width=512,
               transformer_heads=8, transformer_layers=12)
    tower = dict(img_size=224, patch_size=16, vision_layers=12, vision_heads=None)
    for k, v in overrides.items():
        (tower if k in tower else cfg)[k] = v
    torch.manual_seed(seed)
    heads = tower["vision_heads"] or cfg["vision_width"] // 64
    vit = TimmVisionTransformer(tower["img_size"], tower["patch_size"], cfg["vision_width"], tower["vision_layers"], heads)
    model = SlipClip(vision_model=vit, **cfg)
    if trained_like:
        perturb_timm_trained_like(vit, seed + 7001)
        perturb_trained_like(model, seed + 7002)  # CLIP-side names only: the text tower's LayerNorms and biases
    return model.eval()
