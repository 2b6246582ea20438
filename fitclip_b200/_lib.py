"""ctypes binding of ``libfitclip_b200.so`` (declared in ``include/fitclip_b200.h``).

There is deliberately no fallback: if the shared library is missing, or the device is not sm_100, every entry point
raises.  Build it with ``python __graft_entry__.py`` (or ``make -C fitclip_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfitclip_b200.so")
if os.environ.get("FITCLIP_VARIANT"):  # diagnostics: an experiment build of the SAME library (`make VARIANT=name EXTRA=-D...`)
    LIB_PATH = LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_%s.so" % os.environ["FITCLIP_VARIANT"])


class FitclipError(RuntimeError):
    def __init__(self, status: int, message: str) -> None:
        super().__init__(f"libfitclip_b200 error {status}: {message}")
        self.status = status


class fc_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "embed_dim", "image_resolution", "vision_layers", "vision_width", "vision_patch_size", "context_length",
        "vocab_size", "transformer_width", "transformer_heads", "transformer_layers", "max_frames_per_pass",
        "max_texts_per_pass", "vision_tower", "vision_attn_width")]


TOWER_OPENAI, TOWER_TIMM = 0, 1
INTERPOLATION = {"bicubic": 0, "bilinear": 1}


class fc_profile_record(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tag", C.c_int32), ("n", C.c_int64), ("k", C.c_int64),
                ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double),
                ("rows", C.c_double)]


_p, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); must list every FC_API symbol of include/fitclip_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "fc_version": (C.c_int, []),
    "fc_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "fc_launch_count": (_i64, []),
    "fc_profile_start": (C.c_int, [_i32]),
    "fc_profile_stop": (C.c_int, [_p, _i32]),
    "fc_model_create": (C.c_int, [C.POINTER(fc_config), C.POINTER(_p)]),
    "fc_model_destroy": (C.c_int, [_p]),
    "fc_model_set_param": (C.c_int, [_p, C.c_char_p, _p, _i64, _p]),
    "fc_model_ready": (C.c_int, [_p]),
    "fc_model_workspace_bytes": (_i64, [_p]),
    "fc_encode_video": (C.c_int, [_p, _p, C.c_int, _i64, _i32, _p, _p, _p]),
    "fc_encode_video_uint8": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _i32, _p, _p, _p]),
    "fc_encode_text": (C.c_int, [_p, _p, _i64, _p, _p]),
    "fc_model_check": (C.c_int, [_p, _p]),
    "fc_preprocess_frames": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, C.c_int, _i32, _p]),
    "fc_preprocess_to_patches": (C.c_int, [_p, _i64, _i32, _i32, _i32, _i32, _p, _p, _p, _i64, _i32, _p]),
    "fc_pool_normalize": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _f32, _p]),
    "fc_wise_lerp": (C.c_int, [_p, _p, _p, _p, _i64, _f64, _p]),
    "fc_sim_workspace_bytes": (_i64, [_i64, _i64, _i32, _i32]),
    "fc_sim_prepare": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _p, _p]),
    "fc_sim_target_scores": (C.c_int, [_p, _i64, _i64, _i32, _i32, _p, _i32, _p, _p]),
    "fc_sim_count": (C.c_int, [_p, _i64, _i64, _i32, _i32, _p, _i32, _p, _p, _p]),
    "fc_sim_scores": (C.c_int, [_p, _i64, _i64, _i32, _i32, _f32, _p, _i64, _p]),
    "fc_rank_from_scores": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "fc_counts_to_ranks": (C.c_int, [_p, _p, _i64, _p]),
    "fc_metrics_from_ranks": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p]),
    "fc_topk_rows": (C.c_int, [_p, _i64, _i64, _i64, _i32, _p, _p, _p]),
    "fc_nce_loss": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "fc_ts_nce_loss": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p]),
    "fc_gemm_bf16": (C.c_int, [C.c_int, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _f32, _i32, _i32, _i32, _p]),
    "fc_layernorm_bf16": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _f32, _p]),
    "fc_attention_bf16": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p]),
    # training step (row f3)
    "fc_gemm_bf16_splitk": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _f32, _i32, _i32, _i32, _i32, _p]),
    "fc_gemm_bf16_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _f32, _i32,
                                      _i32, _i32, _i32, _p]),
    "fc_colsum_bf16": (C.c_int, [_p, _i64, _i64, _i32, _p, _p]),
    "fc_transpose_bf16": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _p, _p]),
    "fc_layernorm_bwd_bf16": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _f32, _p]),
    "fc_quickgelu_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "fc_quickgelu_bwd_bf16": (C.c_int, [_p, _p, _p, _p, _i64, _p]),
    "fc_attention_bwd_bf16": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "fc_loss_fwd_bwd": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _f32, _p, _p, _i64, _p]),
    "fc_sgemm_f32": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _f32, _p, _i64, _p, _i64, _p, _i64, _p]),
    "fc_pool_normalize_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _f32, _p]),
    "fc_seq_rows": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "fc_seq_sum": (C.c_int, [_p, _p, _i64, _i32, _i32, _p]),
    "fc_token_scatter_add": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _p]),
    "fc_adamw_step": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _p]),
    "fc_f32_to_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "fc_patch_embed": (C.c_int, [_p, C.c_int, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "fc_text_embed": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p]),
}

EPI_BIAS, EPI_BIAS_QGELU, EPI_BIAS_RESID, EPI_F32, EPI_F32_SPLITK, EPI_QGELU_BWD = 0, 1, 2, 4, 9, 11
DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads the shared library (once). Raises if it has not been built -- there is no Python/CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FitclipError(-100, f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                                     f"(or `make -C fitclip_b200/csrc`); there is no fallback path")
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(2048)
    load().fc_last_error(buf, 2048)
    return buf.value.decode("utf-8", "replace")


def check(status: int) -> None:
    if status != 0:
        raise FitclipError(status, last_error())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FitclipError(-101, "expected a CUDA tensor: libfitclip_b200 has no CPU path")
    return t.data_ptr()


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def profile_start(max_records: int = 65536) -> None:
    check(load().fc_profile_start(max_records))


def profile_stop(cap: int = 256):
    """-> list of dict records aggregated by (kind, tag, n, k); synchronises the device."""
    arr = (fc_profile_record * cap)()
    n = load().fc_profile_stop(C.cast(arr, C.c_void_p), cap)
    if n < 0:
        check(n)
    return [{f: getattr(arr[i], f) for f, _ in fc_profile_record._fields_} for i in range(n)]


def launch_count() -> int:
    return int(load().fc_launch_count())
