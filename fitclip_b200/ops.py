"""Tensor-level wrappers over the C ABI (one function per exported kernel group). Everything here runs on the
caller's current CUDA stream and raises :class:`fitclip_b200._lib.FitclipError` on failure -- no fallbacks."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _dev(t: torch.Tensor) -> torch.device:
    if not t.is_cuda:
        raise _lib.FitclipError(-101, "expected a CUDA tensor: libfitclip_b200 has no CPU path")
    return t.device


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None,
              resid: Optional[torch.Tensor] = None, epilogue: int = _lib.EPI_BIAS, alpha: float = 1.0,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``epilogue(a @ b.T)`` with bf16 ``a (M,K)``, ``b (N,K)`` (row strides may exceed K)."""
    dev = _dev(a)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if epilogue == _lib.EPI_F32 else torch.bfloat16)
    with torch.cuda.device(dev):
        check(_lib.load().fc_gemm_bf16(epilogue, ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0),
                                       ptr(bias), ptr(resid), 0 if resid is None else resid.stride(0), alpha, M, N, K,
                                       stream_ptr(dev)))
    return out


def layernorm_bf16(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _dev(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and gamma.dtype == torch.float32
    rows, D = x.reshape(-1, x.shape[-1]).shape
    out = torch.empty_like(x) if out is None else out
    with torch.cuda.device(dev):
        check(_lib.load().fc_layernorm_bf16(ptr(x), ptr(out), ptr(gamma), ptr(beta), rows, D, eps, stream_ptr(dev)))
    return out


def attention_bf16(qkv: torch.Tensor, seqs: int, L: int, heads: int, causal: bool) -> torch.Tensor:
    """``qkv``: bf16 ``(seqs*L, 3*heads*64)`` -> bf16 ``(seqs*L, heads*64)``."""
    dev = _dev(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (seqs * L, 3 * heads * 64)
    out = torch.empty(seqs * L, heads * 64, device=dev, dtype=torch.bfloat16)
    with torch.cuda.device(dev):
        check(_lib.load().fc_attention_bf16(ptr(qkv), ptr(out), seqs, L, heads, int(causal), stream_ptr(dev)))
    return out


def preprocess_frames(frames: torch.Tensor, size: int, mean, std, dtype: torch.dtype = torch.float32,
                      interpolation: str = "bicubic") -> torch.Tensor:
    """uint8 ``(..., H, W, 3)`` frames -> ``(..., 3, size, size)`` of ``dtype`` (fp32 / bf16): the reference's eval
    transform (``aligner/encoder/clip_video_text_encoder.py:124-133``: /255, bicubic resize of the shorter side to
    ``size``, centre crop, normalise) as one kernel on the GPU.  ``interpolation="bilinear"``: the SLIP wrapper's transform
    (``slip_video_text_encoder.py:78-87``)."""
    import ctypes as C
    dev = _dev(frames)
    if frames.dtype != torch.uint8 or frames.dim() < 3 or frames.shape[-1] != 3:
        raise ValueError(f"expected uint8 frames of shape (..., H, W, 3), got {frames.dtype} {tuple(frames.shape)}")
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("preprocess_frames emits fp32 or bf16")
    lead, (H, W) = frames.shape[:-3], frames.shape[-3:-1]
    flat = frames.reshape(-1, H, W, 3).contiguous()
    out = torch.empty(flat.shape[0], 3, size, size, device=dev, dtype=dtype)
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(dev):
        for lo in range(0, flat.shape[0], 65535):  # the kernel takes at most 65535 frames per launch
            chunk = flat[lo:lo + 65535]
            check(_lib.load().fc_preprocess_frames(ptr(chunk), chunk.shape[0], H, W, size, m3, s3, ptr(out[lo:]),
                                                   _lib.DTYPE_CODE[dtype], _lib.INTERPOLATION[interpolation],
                                                   stream_ptr(dev)))
    return out.reshape(*lead, 3, size, size)


def preprocess_to_patches(frames: torch.Tensor, size: int, patch: int, mean, std,
                          interpolation: str = "bicubic") -> torch.Tensor:
    """uint8 ``(n, H, W, 3)`` frames -> the bf16 patch matrix ``(n * (size/patch)**2, 3 * patch * patch)`` of the ViT patch
    embedding (column order ``(c, ky, kx)`` = ``conv1.weight.reshape(width, -1)``): the eval transform of
    :func:`preprocess_frames` fused with the patch gather."""
    import ctypes as C
    dev = _dev(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"expected uint8 frames of shape (n, H, W, 3), got {frames.dtype} {tuple(frames.shape)}")
    n, H, W = frames.shape[:3]
    g = size // patch
    out = torch.empty(n * g * g, 3 * patch * patch, device=dev, dtype=torch.bfloat16)
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(dev):
        check(_lib.load().fc_preprocess_to_patches(ptr(frames.contiguous()), n, H, W, size, patch, m3, s3, ptr(out),
                                                   out.stride(0), _lib.INTERPOLATION[interpolation], stream_ptr(dev)))
    return out


def pool_normalize(x: torch.Tensor, frames_per_row: int, scale: float = 1.0) -> torch.Tensor:
    """``x (B*T, D)`` fp32 -> ``(B, D)``: L2-normalise every row, mean over each group of T rows
    (``aligner/encoder/clip_video_text_encoder.py:85-89``)."""
    dev = _dev(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] % frames_per_row == 0
    B, D = x.shape[0] // frames_per_row, x.shape[1]
    out = torch.empty(B, D, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.load().fc_pool_normalize(ptr(x), ptr(out), None, B, frames_per_row, D, scale, stream_ptr(dev)))
    return out


def wise_lerp(p1: torch.Tensor, p2: torch.Tensor, weight_for_2: float, out: Optional[torch.Tensor] = None,
              out_bf16: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``(1 - w) * p1 + w * p2`` (``aligner/wise.py:16``), bit-exact with torch's evaluation order."""
    dev = _dev(p1)
    assert p1.dtype == torch.float32 and p2.dtype == torch.float32 and p1.shape == p2.shape
    assert p1.is_contiguous() and p2.is_contiguous()
    out = torch.empty_like(p1) if out is None else out
    with torch.cuda.device(dev):
        check(_lib.load().fc_wise_lerp(ptr(p1), ptr(p2), ptr(out), ptr(out_bf16), p1.numel(), float(weight_for_2),
                                       stream_ptr(dev)))
    return out


def rank_from_scores(scores: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """0-based rank of ``target[i]`` in row i (``Rank.update``, ``aligner/metrics.py:16-19``), int64."""
    dev = _dev(scores)
    assert scores.dtype == torch.float32 and scores.dim() == 2 and scores.stride(1) == 1
    rows, cols = scores.shape
    ranks = torch.empty(rows, device=dev, dtype=torch.int64)
    tgt = target.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        check(_lib.load().fc_rank_from_scores(ptr(scores), scores.stride(0), rows, cols, ptr(tgt), ptr(ranks),
                                              stream_ptr(dev)))
    return ranks


def metrics_from_ranks(ranks: torch.Tensor, num_candidates: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (recall@[1,5,10] fp32 (3,), median rank int64 (), mean rank fp32 ())."""
    dev = _dev(ranks)
    assert ranks.dtype == torch.int64 and ranks.is_contiguous()
    recall = torch.empty(3, device=dev, dtype=torch.float32)
    median = torch.empty((), device=dev, dtype=torch.int64)
    mean = torch.empty((), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.load().fc_metrics_from_ranks(ptr(ranks), ranks.numel(), num_candidates, ptr(recall), ptr(median),
                                                ptr(mean), stream_ptr(dev)))
    return recall, median, mean


def topk_rows(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    dev = _dev(scores)
    assert scores.dtype == torch.float32 and scores.dim() == 2 and scores.stride(1) == 1
    rows, cols = scores.shape
    values = torch.empty(rows, k, device=dev, dtype=torch.float32)
    indices = torch.empty(rows, k, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(_lib.load().fc_topk_rows(ptr(scores), scores.stride(0), rows, cols, k, ptr(values), ptr(indices),
                                       stream_ptr(dev)))
    return values, indices


def nce_loss(scores: torch.Tensor) -> torch.Tensor:
    """``nce_loss(scores)`` with mean reduction (``aligner/loss.py:13-26``), forward only."""
    dev = _dev(scores)
    assert scores.dtype == torch.float32 and scores.shape[0] == scores.shape[1] and scores.stride(1) == 1
    B = scores.shape[0]
    ws = torch.empty(2 * B, device=dev, dtype=torch.float32)
    out = torch.empty((), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.load().fc_nce_loss(ptr(scores), scores.stride(0), B, ptr(ws), ptr(out), stream_ptr(dev)))
    return out


def teacher_student_nce_loss(scores: torch.Tensor, teacher_scores: torch.Tensor) -> torch.Tensor:
    """``TeacherStudentNCELoss(reduction="batchmean")`` (``aligner/loss.py:29-39``, ``teacher_student.py:73``)."""
    dev = _dev(scores)
    assert scores.shape == teacher_scores.shape and scores.shape[0] == scores.shape[1]
    assert scores.dtype == torch.float32 and teacher_scores.dtype == torch.float32
    assert scores.stride(1) == 1 and teacher_scores.stride() == scores.stride()
    B = scores.shape[0]
    ws = torch.empty(2 * B, device=dev, dtype=torch.float32)
    out = torch.empty((), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.load().fc_ts_nce_loss(ptr(scores), ptr(teacher_scores), scores.stride(0), B, ptr(ws), ptr(out),
                                         stream_ptr(dev)))
    return out


class Similarity:
    """Prepared operands of ``texts @ videos.T`` for one (possibly sharded) column slab.

    ``terms=3`` splits every fp32 embedding into bf16 hi/lo parts so that the tensor-core product carries ~16 mantissa
    bits; ``terms=1`` rounds the operands to bf16 once."""

    def __init__(self, text_emb: torch.Tensor, video_emb: torch.Tensor, terms: int = 3) -> None:
        dev = _dev(text_emb)
        assert text_emb.dtype == torch.float32 and video_emb.dtype == torch.float32
        assert text_emb.is_contiguous() and video_emb.is_contiguous() and text_emb.shape[1] == video_emb.shape[1]
        self.nt, self.dim = text_emb.shape
        self.nv = video_emb.shape[0]
        self.terms = terms
        self.device = dev
        lib = _lib.load()
        self.ws = torch.empty(lib.fc_sim_workspace_bytes(self.nt, self.nv, self.dim, terms), device=dev,
                              dtype=torch.uint8)
        with torch.cuda.device(dev):
            check(lib.fc_sim_prepare(ptr(text_emb), ptr(video_emb), self.nt, self.nv, self.dim, terms, ptr(self.ws),
                                     stream_ptr(dev)))

    def scores(self, alpha: float = 1.0) -> torch.Tensor:
        out = torch.empty(self.nt, self.nv, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(_lib.load().fc_sim_scores(ptr(self.ws), self.nt, self.nv, self.dim, self.terms, alpha, ptr(out),
                                            out.stride(0), stream_ptr(self.device)))
        return out

    def target_scores(self, target: torch.Tensor, col_offset: int = 0) -> torch.Tensor:
        """score of each row's target column where that column is in this slab, 0 elsewhere."""
        out = torch.zeros(self.nt, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(_lib.load().fc_sim_target_scores(ptr(self.ws), self.nt, self.nv, self.dim, self.terms, ptr(target),
                                                   col_offset, ptr(out), stream_ptr(self.device)))
        return out

    def counts(self, target: torch.Tensor, target_scores: torch.Tensor, col_offset: int = 0) -> torch.Tensor:
        """int32 per row: #columns of this slab that outrank the target (tie rule in ``include/fitclip_b200.h``)."""
        out = torch.zeros(self.nt, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            check(_lib.load().fc_sim_count(ptr(self.ws), self.nt, self.nv, self.dim, self.terms, ptr(target),
                                           col_offset, ptr(target_scores), ptr(out), stream_ptr(self.device)))
        return out
