"""Drop-in for the reference's CLIP encoder plugin on B200.

* :class:`B200Clip` stands where ``clip.model.CLIP`` stands in the reference: an ``nn.Module`` whose parameters carry
  the OpenAI state-dict names (``visual.conv1.weight`` ... ``text_projection``) and which offers ``encode_image`` /
  ``encode_text`` / ``visual.input_resolution`` -- the three things ``ClipVideoTextEncoder`` touches
  (``aligner/encoder/clip_video_text_encoder.py:84,93,115``).  The arithmetic runs in ``libfitclip_b200.so``.
* :class:`B200ClipVideoTextEncoder` mirrors ``ClipVideoTextEncoder`` (``clip_video_text_encoder.py:68-146``): same
  constructor, hooks, normalisation and pooling semantics.
* :func:`load_clip_model` mirrors ``load_clip_model`` (``:30-61``) for local state-dict files and modules.

No CPU path: encoding CPU tensors, or running without the built extension / an sm_100 device, raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, Iterable, Iterator, Mapping, Optional, Union

import torch
from torch import nn

from . import _lib, tokenizer
from .api import TYPE_TEXT_INPUT, TYPE_TOKENIZER, TYPE_TRANSFORM, TYPE_VIDEO_INPUT, VideoTextEncoder, \
    float_standard_denormalize
from .frame_sampler import RandomFromUniformIntervalsFrameSampler, UniformFrameSampler

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # clip_video_text_encoder.py:72
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def infer_config(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, int]:
    """Geometry from tensor shapes, the way ``clip.model.build_model`` does it."""
    if any(k.startswith("visual.layer1.") for k in state_dict):
        # config/encoder/clip_rn50.yaml ... clip_rn50x64.yaml, open_clip_rn*.yaml: CLIP's ModifiedResNet image tower
        raise _lib.FitclipError(-1, "this state dict holds a ModifiedResNet image tower (CLIP RN50 / RN101 / RN50x*): only "
                                    "the VisionTransformer towers are built (ViT-B/32, B/16, L/14, L/14@336, SLIP layout)")
    if "visual.pos_embed" in state_dict or "module.visual.pos_embed" in state_dict:
        raise _lib.FitclipError(-1, "this is a SLIP-layout state dict (timm image tower): use B200SlipClip / "
                                    "B200SlipVideoTextEncoder / load_slip_model")
    conv = state_dict["visual.conv1.weight"]
    vision_width, patch = conv.shape[0], conv.shape[-1]
    grid = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    width = state_dict["ln_final.weight"].shape[0]
    return dict(
        embed_dim=state_dict["text_projection"].shape[1], image_resolution=patch * grid,
        vision_layers=len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")]),
        vision_width=vision_width, vision_patch_size=patch, context_length=state_dict["positional_embedding"].shape[0],
        vocab_size=state_dict["token_embedding.weight"].shape[0], transformer_width=width,
        transformer_heads=width // 64,
        transformer_layers=len({k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")}))


def _version_of(t: torch.Tensor) -> int:
    """In-place edit counter; tensors created under ``torch.inference_mode()`` have none (they cannot be edited in place
    outside inference mode either), so their storage pointer alone identifies them."""
    try:
        return t._version
    except RuntimeError:
        return -1


class _Node(nn.Module):
    """Anonymous container used to reproduce the dotted OpenAI parameter names."""


class _Engine:
    """Owns the native ``fc_model`` handle. Never copied or pickled: a copy starts without a handle and re-creates
    it lazily (``aligner/wise.py:21`` deep-copies encoders)."""

    def __init__(self) -> None:
        self.handle: Optional[C.c_void_p] = None
        self.device: Optional[torch.device] = None
        self.signature: Any = None

    def __deepcopy__(self, memo) -> "_Engine":
        return _Engine()

    def __getstate__(self):
        return {}

    def __setstate__(self, state) -> None:
        self.__init__()

    def close(self) -> None:
        if self.handle is not None:
            try:
                _lib.load().fc_model_destroy(self.handle)
            finally:
                self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


class B200Clip(nn.Module):
    """OpenAI-layout CLIP parameters + the native forward. ``source`` is a state dict or any module whose
    ``state_dict()`` uses the OpenAI names (e.g. ``clip.model.CLIP``); ``logit_scale`` is kept if present so that
    checkpoints load strictly, exactly like the reference's model object."""

    def __init__(self, source: Union[Mapping[str, torch.Tensor], nn.Module], max_frames_per_pass: Optional[int] = None,
                 max_texts_per_pass: int = 1024) -> None:
        super().__init__()
        state_dict = source.state_dict() if isinstance(source, nn.Module) else source
        state_dict = self._filter_state_dict(state_dict)
        self.config = self._infer_config(state_dict)
        if max_frames_per_pass is None:
            # ~400k image tokens per pass (ViT-B/16: 2030 frames, ViT-L/14: 1556, ViT-L/14@336: 693): measured on the
            # bench shape, 1920-frame passes run 1 % faster than 500-frame ones (fewer launches and wave tails; L2
            # residency of the activations does not matter); the workspace grows to ~0.9 GB per 100k tokens of ViT-B/16
            tokens = (self.config["image_resolution"] // self.config["vision_patch_size"]) ** 2 + 1
            max_frames_per_pass = max(64, 400_000 // tokens)
        self.max_frames_per_pass = max_frames_per_pass
        self.max_texts_per_pass = max_texts_per_pass
        for name, value in state_dict.items():
            *path, leaf = name.split(".")
            node: nn.Module = self
            for part in path:
                if not hasattr(node, part):
                    node.add_module(part, _Node())
                node = getattr(node, part)
            # fp32 like load_clip_in_float32 (clip_video_text_encoder.py:22-25)
            node.register_parameter(leaf, nn.Parameter(value.detach().clone().float(), requires_grad=False))
        self.visual.input_resolution = self.config["image_resolution"]
        self.visual.output_dim = self.config["embed_dim"]
        self.context_length = self.config["context_length"]
        self.vocab_size = self.config["vocab_size"]
        self._engine = _Engine()

    # ---- layout hooks (overridden by the SLIP layout, slip_encoder.py) ----------------------------------------------
    @staticmethod
    def _filter_state_dict(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}

    @staticmethod
    def _infer_config(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, int]:
        return infer_config(state_dict)

    def _engine_params(self):
        """``(name the engine knows, fp32 tensor)`` for every weight the native model needs."""
        return [(n, p) for n, p in self.named_parameters() if n != "logit_scale"]

    @property
    def dtype(self) -> torch.dtype:
        return self.text_projection.dtype

    # ---- native engine ------------------------------------------------------------------------------------------
    def _native(self, device: torch.device) -> C.c_void_p:
        """Handle with weights in sync with the current nn.Parameters (re-uploaded after ``load_state_dict``,
        ``.to()``, WiSE, in-place edits: detected through tensor versions / storage pointers)."""
        params = [(n, p) for n, p in self.named_parameters() if n != "logit_scale"]
        for _, p in params:
            if p.device != device or p.dtype != torch.float32:
                raise _lib.FitclipError(-101, f"B200Clip parameters must be fp32 on {device} (found {p.dtype} on "
                                              f"{p.device}); call .to(device).float() first")
        signature = tuple((p.data_ptr(), _version_of(p)) for _, p in params)
        eng, lib = self._engine, _lib.load()
        with torch.cuda.device(device):
            if eng.handle is None or eng.device != device:
                eng.close()
                cfg = _lib.fc_config(**self.config, max_frames_per_pass=self.max_frames_per_pass,
                                     max_texts_per_pass=self.max_texts_per_pass)
                handle = C.c_void_p()
                _lib.check(lib.fc_model_create(C.byref(cfg), C.byref(handle)))
                eng.handle, eng.device, eng.signature = handle, device, None
            if eng.signature != signature:
                stream = _lib.stream_ptr(device)
                for name, p in self._engine_params():
                    data = p.detach().contiguous()
                    _lib.check(lib.fc_model_set_param(eng.handle, name.encode(), data.data_ptr(), data.numel(), stream))
                if not lib.fc_model_ready(eng.handle):
                    raise _lib.FitclipError(-4, _lib.last_error())
                eng.signature = signature
        return eng.handle

    def encode_video_pooled(self, video: torch.Tensor) -> torch.Tensor:
        """(B, T, 3, R, R) -> (B, E): encode_image + per-frame L2 normalise + mean over T, one native call."""
        if not video.is_cuda:
            raise _lib.FitclipError(-101, "B200Clip needs CUDA tensors: there is no CPU path")
        if video.dtype not in _lib.DTYPE_CODE:
            video = video.float()
        B, T = video.shape[:2]
        R = self.config["image_resolution"]
        if tuple(video.shape[2:]) != (3, R, R):
            raise ValueError(f"expected frames of shape (3, {R}, {R}), got {tuple(video.shape[2:])}")
        video = video.contiguous()
        out = torch.empty(B, self.config["embed_dim"], device=video.device, dtype=torch.float32)
        handle = self._native(video.device)
        with torch.cuda.device(video.device):
            _lib.check(_lib.load().fc_encode_video(handle, _lib.ptr(video), _lib.DTYPE_CODE[video.dtype], B, T,
                                                   _lib.ptr(out), None, _lib.stream_ptr(video.device)))
        return out

    def encode_video_uint8_pooled(self, video: torch.Tensor, mean, std, interpolation: str = "bicubic") -> torch.Tensor:
        """raw frames (B, T, H, W, 3) uint8 -> (B, E): eval transform fused into the patch gather + the encoder, one
        native call (``fc_encode_video_uint8``)."""
        if not video.is_cuda:
            raise _lib.FitclipError(-101, "B200Clip needs CUDA tensors: there is no CPU path")
        if video.dtype != torch.uint8 or video.dim() != 5 or video.shape[-1] != 3:
            raise ValueError(f"expected uint8 frames of shape (B, T, H, W, 3), got {video.dtype} {tuple(video.shape)}")
        B, T, H, W = video.shape[:4]
        video = video.contiguous()
        out = torch.empty(B, self.config["embed_dim"], device=video.device, dtype=torch.float32)
        handle = self._native(video.device)
        m3 = (C.c_float * 3)(*[float(v) for v in mean])
        s3 = (C.c_float * 3)(*[float(v) for v in std])
        with torch.cuda.device(video.device):
            _lib.check(_lib.load().fc_encode_video_uint8(handle, _lib.ptr(video), B, T, H, W, m3, s3,
                                                         _lib.INTERPOLATION[interpolation], _lib.ptr(out), None,
                                                         _lib.stream_ptr(video.device)))
        return out

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        """``CLIP.encode_image``: (F, 3, R, R) -> un-normalised (F, E) fp32."""
        if not image.is_cuda:
            raise _lib.FitclipError(-101, "B200Clip needs CUDA tensors: there is no CPU path")
        if image.dtype not in _lib.DTYPE_CODE:
            image = image.float()
        image = image.contiguous()
        F = image.shape[0]
        E = self.config["embed_dim"]
        pooled = torch.empty(F, E, device=image.device, dtype=torch.float32)
        feats = torch.empty(F, E, device=image.device, dtype=torch.float32)
        handle = self._native(image.device)
        with torch.cuda.device(image.device):
            _lib.check(_lib.load().fc_encode_video(handle, _lib.ptr(image), _lib.DTYPE_CODE[image.dtype], F, 1,
                                                   _lib.ptr(pooled), _lib.ptr(feats), _lib.stream_ptr(image.device)))
        return feats

    def encode_text_normalized(self, text: torch.Tensor) -> torch.Tensor:
        """token ids (C, context_length) -> L2-normalised (C, E) fp32."""
        if not text.is_cuda:
            raise _lib.FitclipError(-101, "B200Clip needs CUDA tensors: there is no CPU path")
        if text.dim() != 2 or text.shape[1] != self.context_length:
            raise ValueError(f"expected token ids of shape (C, {self.context_length}), got {tuple(text.shape)}")
        ids = text.to(torch.int32).contiguous()
        out = torch.empty(ids.shape[0], self.config["embed_dim"], device=ids.device, dtype=torch.float32)
        handle = self._native(ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(_lib.load().fc_encode_text(handle, _lib.ptr(ids), ids.shape[0], _lib.ptr(out),
                                                  _lib.stream_ptr(ids.device)))
        return out

    def check_inputs(self) -> None:
        """Synchronises and raises if a token id was out of range (torch's embedding lookup would have raised)."""
        eng = self._engine
        if eng.handle is not None:
            with torch.cuda.device(eng.device):
                _lib.check(_lib.load().fc_model_check(eng.handle, _lib.stream_ptr(eng.device)))


def load_clip_model(name: Union[str, os.PathLike, Mapping[str, torch.Tensor], nn.Module], *args,
                    **kwargs) -> B200Clip:
    """``load_clip_model`` (``clip_video_text_encoder.py:30-61``) for what exists offline: a path to a ``torch.save``d
    OpenAI-layout state dict (``logit_scale`` optional, ``:45-53``), a state dict, or a module.  Model names / URLs
    need the network and the ``clip`` package and raise a clear error here."""
    if isinstance(name, (str, os.PathLike)):
        if not os.path.exists(name):
            raise FileNotFoundError(f"{name!r}: pretrained CLIP names/URLs cannot be resolved offline; pass a local "
                                    f"state-dict file")
        name = torch.load(name, map_location="cpu")
        if isinstance(name, Mapping) and "state_dict" in name:
            name = name["state_dict"]
    device = kwargs.pop("device", args[0] if args else "cpu")  # reference default: cpu (:55-56)
    return B200Clip(name, **kwargs).to(device)


class B200ClipVideoTextEncoder(VideoTextEncoder):
    """``ClipVideoTextEncoder`` on libfitclip_b200. ``model`` may be a :class:`B200Clip`, any OpenAI-layout module
    (``clip.model.CLIP``, the oracle's restatement) or a state dict."""

    def __init__(self, model: Union[B200Clip, nn.Module, Mapping[str, torch.Tensor]], num_frames: int = 4) -> None:
        super().__init__()
        self.model = model if isinstance(model, B200Clip) else B200Clip(model)
        self.num_frames = num_frames
        # The reference unregisters logit_scale (clip_video_text_encoder.py:75-77).
        if hasattr(self.model, "logit_scale"):
            delattr(self.model, "logit_scale")

    def encode_video(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        # clip_video_text_encoder.py:80-89 -- fused natively: encode_image, x/||x|| per frame, mean over frames
        return self.model.encode_video_pooled(video)

    def encode_video_uint8(self, video: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Raw decoded frames ``(B, T, H, W, 3)`` uint8 on the GPU -> ``(B, E)``: the eval transform
        (``get_eval_transform``, clip_video_text_encoder.py:124-133) runs on the GPU in front of the encoder instead of
        on DataLoader workers.  Default (``dtype=None``): fused -- the transform writes bf16 patch rows straight into
        the patch-embedding GEMM's operand, no normalised frame is ever stored.  ``dtype=torch.float32 / bfloat16``:
        the two-step path through an NCHW intermediate of that precision (``fc_preprocess_frames``)."""
        if dtype is None:
            return self.model.encode_video_uint8_pooled(video, CLIP_MEAN, CLIP_STD)
        from . import ops
        frames = ops.preprocess_frames(video, self.model.visual.input_resolution, CLIP_MEAN, CLIP_STD, dtype)
        return self.model.encode_video_pooled(frames)

    def encode_text(self, text: TYPE_TEXT_INPUT) -> torch.Tensor:
        # clip_video_text_encoder.py:92-94
        return self.model.encode_text_normalized(text["input_ids"])

    def get_tokenizer(self) -> TYPE_TOKENIZER:
        return tokenizer.tokenize

    def decode_text(self, text: TYPE_TEXT_INPUT) -> Iterator[str]:
        return tokenizer.decode(text["input_ids"] if isinstance(text, Mapping) else (t["input_ids"] for t in text))

    def get_train_frame_sampler(self):
        return RandomFromUniformIntervalsFrameSampler(self.num_frames)

    def get_eval_frame_sampler(self):
        return UniformFrameSampler(self.num_frames)

    def get_train_transform(self, dtype: torch.dtype) -> TYPE_TRANSFORM:
        from .transforms import train_transform
        return train_transform(self.model.visual.input_resolution, dtype, CLIP_MEAN, CLIP_STD)

    def get_eval_transform(self, dtype: torch.dtype) -> TYPE_TRANSFORM:
        from .transforms import eval_transform
        return eval_transform(self.model.visual.input_resolution, dtype, CLIP_MEAN, CLIP_STD)

    @property
    def should_pad_batch(self) -> bool:
        return True

    def to_bchw(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def denormalize_video_tensor(self, video: TYPE_VIDEO_INPUT) -> torch.Tensor:
        return float_standard_denormalize(video, mean=CLIP_MEAN, std=CLIP_STD)
