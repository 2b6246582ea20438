"""SLIP-layout models on libfitclip_b200.

The reference evaluates the SLIP repository's checkpoints (``config/encoder/slip*.yaml``, ``clip_vit_b_16_cc12m.yaml`` ...)
through ``SlipVideoTextEncoder`` (``aligner/encoder/slip_video_text_encoder.py:26-100``) around ``slip.CLIP`` / ``slip.SLIP``
(``aligner/encoder/slip.py:399-480,530-557``): a timm ``VisionTransformer`` (``vit_base_patch16_224`` /
``vit_large_patch16_224``, ``num_classes=0``) + ``image_projection`` and CLIP's own text tower.  Pooling is the CLIP
wrapper's (``:37-51``): per-frame L2 normalisation, mean over the frames, no re-normalisation.

* :class:`B200SlipClip` keeps the parameters under the checkpoint's own names (``visual.blocks.3.attn.qkv.weight``,
  ``visual.pos_embed``, ``image_projection`` ...), so SLIP state dicts load strictly and ``wise`` lerps two of them name by
  name; the native engine receives them under the OpenAI names it knows (``FC_TOWER_TIMM``: no ``ln_pre``, exact GELU,
  LayerNorm eps 1e-6), the patch-embedding bias folded into the positional rows of the patch tokens.
* :class:`B200SlipVideoTextEncoder` mirrors ``SlipVideoTextEncoder``: same hooks, ImageNet statistics, bilinear resize.
* :func:`load_slip_model` mirrors ``load_model`` (``:19-23``) for local checkpoint files.

Heads narrower than 64 (the ViT-S/16 variants: ``vit_small_mocov3_patch16_224``, 384 wide, 12 heads of 32, ``slip.py:566-569``)
run with every head in a 64-wide slot of zero-padded q / k / v rows and out_proj columns (``fc_config.vision_attn_width``).
"""
from __future__ import annotations

import os
import re
from typing import Dict, Iterator, Mapping, Optional, Union

import torch
from torch import nn

from . import _lib, tokenizer
from .api import TYPE_TEXT_INPUT, TYPE_TOKENIZER, TYPE_TRANSFORM, TYPE_VIDEO_INPUT, VideoTextEncoder, \
    float_standard_denormalize
from .encoder import B200Clip
from .frame_sampler import UniformFrameSampler

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # slip_video_text_encoder.py:91
IMAGENET_STD = (0.229, 0.224, 0.225)

_BLOCK = re.compile(r"^visual\.blocks\.(\d+)\.(.+)$")
_BLOCK_LEAF = {
    "norm1.weight": "ln_1.weight", "norm1.bias": "ln_1.bias", "norm2.weight": "ln_2.weight", "norm2.bias": "ln_2.bias",
    "attn.qkv.weight": "attn.in_proj_weight", "attn.qkv.bias": "attn.in_proj_bias",
    "attn.proj.weight": "attn.out_proj.weight", "attn.proj.bias": "attn.out_proj.bias",
    "mlp.fc1.weight": "mlp.c_fc.weight", "mlp.fc1.bias": "mlp.c_fc.bias",
    "mlp.fc2.weight": "mlp.c_proj.weight", "mlp.fc2.bias": "mlp.c_proj.bias",
}
_TOP = {"visual.patch_embed.proj.weight": "visual.conv1.weight", "visual.norm.weight": "visual.ln_post.weight",
        "visual.norm.bias": "visual.ln_post.bias", "image_projection": "visual.proj"}


def is_slip_layout(state_dict: Mapping[str, torch.Tensor]) -> bool:
    return any(k.endswith("visual.pos_embed") for k in state_dict)


def infer_slip_config(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, int]:
    """Geometry from tensor shapes (the constructors of ``slip.py:571-640`` fix everything but the tower size)."""
    conv = state_dict["visual.patch_embed.proj.weight"]
    vision_width, patch = conv.shape[0], conv.shape[-1]
    grid = round((state_dict["visual.pos_embed"].shape[1] - 1) ** 0.5)
    width = state_dict["ln_final.weight"].shape[0]
    return dict(
        embed_dim=state_dict["text_projection"].shape[1], image_resolution=patch * grid,
        vision_layers=len([k for k in state_dict if _BLOCK.match(k) and k.endswith(".attn.qkv.weight")]),
        vision_width=vision_width, vision_patch_size=patch, context_length=state_dict["positional_embedding"].shape[0],
        vocab_size=state_dict["token_embedding.weight"].shape[0], transformer_width=width, transformer_heads=width // 64,
        transformer_layers=len({k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")}),
        vision_tower=_lib.TOWER_TIMM)


class B200SlipClip(B200Clip):
    """``slip.CLIP`` / ``slip.SLIP`` parameters (their own names) + the native forward.  ``vision_heads`` defaults to
    ``vision_width / 64`` (ViT-B/16: 12, ViT-L/16: 16); any other head dimension raises."""

    def __init__(self, source: Union[Mapping[str, torch.Tensor], nn.Module], vision_heads: Optional[int] = None,
                 **kwargs) -> None:
        super().__init__(source, **kwargs)
        width = self.config["vision_width"]
        if vision_heads is None and width == 384:
            # the SLIP repository's ViT-S/16 is the MoCo-v3 variant with 12 heads of 32 (slip.py:566-569), timm's stock
            # vit_small_patch16_224 has 6 heads of 64: the state dict cannot tell them apart
            raise _lib.FitclipError(-1, "384-wide SLIP-layout vision tower: pass vision_heads (12 for the SLIP repository's "
                                        "ViT-S/16 checkpoints, 6 for timm's stock vit_small_patch16_224)")
        self.vision_heads = vision_heads or width // 64
        head_dim, rem = divmod(width, self.vision_heads)
        if rem or head_dim > 64 or 64 % head_dim:
            raise _lib.FitclipError(-1, f"SLIP-layout vision tower with {self.vision_heads} heads over width {width}: the "
                                        f"attention kernels take head dimensions that divide 64")
        self.vision_head_dim = head_dim
        if head_dim < 64:
            # every head gets a 64-wide slot (zero rows / columns): include/fitclip_b200.h, fc_config.vision_attn_width
            if self.vision_heads * 64 > 1024:
                raise _lib.FitclipError(-1, f"{self.vision_heads} heads padded to 64 exceed the 1024-wide attention limit")
            self.config["vision_attn_width"] = self.vision_heads * 64

    def _pad_heads(self, t: torch.Tensor, groups: int, dim: int) -> torch.Tensor:
        """``t`` holds ``groups * heads * head_dim`` entries along ``dim`` (q | k | v blocks of in_proj, or the input
        columns of out_proj): spread every head over a 64-wide slot, zeros in the upper part."""
        hd, heads = self.vision_head_dim, self.vision_heads
        shape = list(t.shape)
        shape[dim:dim + 1] = [groups, heads, hd]
        t = t.reshape(shape)
        pad = [0, 0] * (t.dim() - (dim + 3)) + [0, 64 - hd]  # F.pad counts dimensions from the last
        t = torch.nn.functional.pad(t, pad)
        shape[dim + 2] = 64
        out = list(t.shape)
        out[dim:dim + 3] = [groups * heads * 64]
        return t.reshape(out)

    @staticmethod
    def _filter_state_dict(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        # "module." is DistributedDataParallel's prefix (slip_video_text_encoder.py:22); image_mlp.* are SLIP's SSL heads,
        # which encode_image never touches (slip.py:530-557)
        out = {}
        for k, v in state_dict.items():
            k = k[len("module."):] if k.startswith("module.") else k
            if not k.startswith("image_mlp."):
                out[k] = v
        return out

    @staticmethod
    def _infer_config(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, int]:
        return infer_slip_config(state_dict)

    def _engine_params(self):
        out = []
        for name, p in self.named_parameters():
            if name == "logit_scale":
                continue
            m = _BLOCK.match(name)
            if m:
                leaf = m.group(2)
                if self.vision_head_dim < 64 and leaf.startswith("attn."):
                    if leaf.startswith("attn.qkv."):
                        # the kernels' softmax scale is 64^-0.5: the q rows carry the rest of head_dim^-0.5
                        p = p.detach().clone()
                        p[:self.config["vision_width"]] *= (64 / self.vision_head_dim) ** 0.5
                        p = self._pad_heads(p, 3, 0)
                    elif leaf == "attn.proj.weight":
                        p = self._pad_heads(p.detach(), 1, 1)
                out.append((f"visual.transformer.resblocks.{m.group(1)}.{_BLOCK_LEAF[leaf]}", p))
            elif name in _TOP:
                out.append((_TOP[name], p))
            elif name == "visual.cls_token":
                out.append(("visual.class_embedding", p.reshape(-1)))
            elif name == "visual.pos_embed":
                # x = conv(patch) + bias + pos[1 + i]: the bias rides on the positional rows of the patch tokens
                pos = p[0].clone()
                pos[1:] += self.visual.patch_embed.proj.bias
                out.append(("visual.positional_embedding", pos))
            elif name == "visual.patch_embed.proj.bias":
                continue
            elif name.startswith("visual."):
                raise _lib.FitclipError(-1, f"unexpected SLIP-layout vision parameter {name!r}")
            else:
                out.append((name, p))
        return out


def load_slip_model(path: Union[str, os.PathLike, Mapping], **kwargs) -> B200SlipClip:
    """``load_model`` (slip_video_text_encoder.py:19-23) for a local checkpoint file or an already loaded checkpoint:
    ``{"args": Namespace(model="CLIP_VITB16" | "SLIP_VITB16" | ...), "state_dict": {"module.visual...": ...}}``, or a bare
    state dict.  URLs need the network and raise."""
    if isinstance(path, (str, os.PathLike)):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path!r}: SLIP checkpoints cannot be downloaded offline; pass a local file")
        path = torch.load(path, map_location="cpu", weights_only=False)
    vision_heads = kwargs.pop("vision_heads", None)
    if isinstance(path, Mapping) and "state_dict" in path:
        arch = getattr(path.get("args"), "model", None)
        if arch is not None and arch.upper().endswith("VITS16"):
            vision_heads = 12  # vit_small_mocov3_patch16_224: 384 wide, 12 heads of 32 (slip.py:566-569)
        path = path["state_dict"]
    device = kwargs.pop("device", "cpu")
    return B200SlipClip(path, vision_heads=vision_heads, **kwargs).to(device)


class B200SlipVideoTextEncoder(VideoTextEncoder):
    """``SlipVideoTextEncoder`` on libfitclip_b200.  ``model``: a :class:`B200SlipClip`, a SLIP-layout module or state dict."""

    def __init__(self, model: Union[B200SlipClip, nn.Module, Mapping[str, torch.Tensor]], num_frames: int = 4) -> None:
        super().__init__()
        self.model = model if isinstance(model, B200SlipClip) else B200SlipClip(model)
        self.num_frames = num_frames
        if hasattr(self.model, "logit_scale"):  # slip_video_text_encoder.py:33-35
            delattr(self.model, "logit_scale")

    def encode_video(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        # slip_video_text_encoder.py:37-47 -- fused natively: encode_image, x/||x|| per frame, mean over frames
        return self.model.encode_video_pooled(video)

    def encode_video_uint8(self, video: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Raw decoded frames ``(B, T, H, W, 3)`` uint8 on the GPU -> ``(B, E)``: ``get_eval_transform`` (bilinear resize,
        centre crop, ImageNet statistics; slip_video_text_encoder.py:78-87) on the GPU, fused into the patch gather by
        default (``dtype=None``) or through an NCHW intermediate of ``dtype``."""
        if dtype is None:
            return self.model.encode_video_uint8_pooled(video, IMAGENET_MEAN, IMAGENET_STD, "bilinear")
        from . import ops
        frames = ops.preprocess_frames(video, self.model.visual.input_resolution, IMAGENET_MEAN, IMAGENET_STD, dtype,
                                       "bilinear")
        return self.model.encode_video_pooled(frames)

    def encode_text(self, text: TYPE_TEXT_INPUT) -> torch.Tensor:
        # slip_video_text_encoder.py:49-51
        return self.model.encode_text_normalized(text["input_ids"])

    def get_tokenizer(self) -> TYPE_TOKENIZER:
        return tokenizer.slip_tokenize

    def decode_text(self, text: TYPE_TEXT_INPUT) -> Iterator[str]:
        return tokenizer.decode(text["input_ids"] if isinstance(text, Mapping) else (t["input_ids"] for t in text))

    def get_train_frame_sampler(self):
        raise NotImplementedError  # slip_video_text_encoder.py:66-67

    def get_eval_frame_sampler(self):
        return UniformFrameSampler(self.num_frames)

    def get_train_transform(self, dtype: torch.dtype) -> TYPE_TRANSFORM:
        raise NotImplementedError  # slip_video_text_encoder.py:74-75

    def get_eval_transform(self, dtype: torch.dtype) -> TYPE_TRANSFORM:
        # slip_video_text_encoder.py:78-87: Resize's default interpolation (bilinear), ImageNet statistics
        from torchvision.transforms import InterpolationMode
        from .transforms import eval_transform
        return eval_transform(self.model.visual.input_resolution, dtype, IMAGENET_MEAN, IMAGENET_STD,
                              interpolation=InterpolationMode.BILINEAR)

    @property
    def should_pad_batch(self) -> bool:
        return True

    def to_bchw(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def denormalize_video_tensor(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        return float_standard_denormalize(video, mean=IMAGENET_MEAN, std=IMAGENET_STD)
