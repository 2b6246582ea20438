"""Tensor-level wrappers over the training-step entry points of the C ABI (``include/fitclip_b200.h``, "training step"
section; kernels in ``fitclip_b200/csrc/train.cu``).  Same rules as :mod:`fitclip_b200.ops`: CUDA tensors only, the
caller's current stream, :class:`FitclipError` on failure, no fallback."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .ops import _dev

BF16 = torch.bfloat16


def _call(dev: torch.device, name: str, *args) -> None:
    with torch.cuda.device(dev):
        check(getattr(_lib.load(), name)(*args, stream_ptr(dev)))


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def gemm_splitk(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, alpha: float = 1.0, k: Optional[int] = None,
                k_splits: int = 0) -> torch.Tensor:
    """``out (M,N) fp32 += alpha * a[:, :k] @ b[:, :k].T`` (bf16 operands whose row pitch may exceed ``k``)."""
    dev = _dev(a)
    assert a.dtype == BF16 and b.dtype == BF16 and out.dtype == torch.float32
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    k = a.shape[1] if k is None else k
    _call(dev, "fc_gemm_bf16_splitk", ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), alpha,
          a.shape[0], b.shape[0], k, k_splits)
    return out


def gemm_nt(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], resid: Optional[torch.Tensor] = None,
            f32: bool = False, qgelu_bwd_of: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``a (M,K) @ w (K,N)`` with ``w`` read in place as an MN-major B operand: the dgrad ``dX = dY @ W`` with ``W`` stored
    ``(N_w, K_w)`` as the forward pass has it (``K`` here = ``N_w``).  bf16 out (+ bias [+ resid]) or fp32 out.
    ``qgelu_bwd_of=u``: bf16 out = ``(a @ w) * quickgelu'(u)`` -- the dgrad of ``c_proj`` and QuickGELU's backward in one
    kernel (``u (M, N)``: the pre-activation the forward kept)."""
    dev = _dev(a)
    assert a.dtype == BF16 and w.dtype == BF16 and a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[0]
    M, K = a.shape
    N = w.shape[1]
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else BF16)
    epi = _lib.EPI_F32 if f32 else (_lib.EPI_BIAS if resid is None else _lib.EPI_BIAS_RESID)
    if qgelu_bwd_of is not None:
        assert not f32 and resid is None and qgelu_bwd_of.dtype == BF16 and qgelu_bwd_of.shape == (M, N)
        assert qgelu_bwd_of.stride(1) == 1
        epi, resid, bias = _lib.EPI_QGELU_BWD, qgelu_bwd_of, None
    _call(dev, "fc_gemm_bf16_layout", epi, 0, 1, ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0),
          ptr(bias), ptr(resid), 0 if resid is None else resid.stride(0), 1.0, M, N, K, 1)
    return out


def wgrad_tn(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, alpha: float = 1.0, k_splits: int = 0,
             rows: Optional[int] = None) -> torch.Tensor:
    """``out (N,K) fp32 += alpha * dy[:rows].T @ x[:rows]`` with ``dy (rows,N)`` and ``x (rows,K)`` read in place (both
    operands MN-major), the token dimension split across the SMs."""
    dev = _dev(dy)
    assert dy.dtype == BF16 and x.dtype == BF16 and out.dtype == torch.float32
    assert dy.stride(1) == 1 and x.stride(1) == 1 and out.stride(1) == 1
    rows = dy.shape[0] if rows is None else rows
    assert x.shape[0] >= rows and out.shape == (dy.shape[1], x.shape[1])
    _call(dev, "fc_gemm_bf16_layout", _lib.EPI_F32_SPLITK, 1, 1, ptr(dy), dy.stride(0), ptr(x), x.stride(0), ptr(out),
          out.stride(0), None, None, 0, alpha, dy.shape[1], x.shape[1], rows, k_splits)
    return out


def colsum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``out (cols) fp32 += x.sum(0)`` for bf16 ``x (rows, cols)``."""
    dev = _dev(x)
    assert x.dtype == BF16 and x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32
    assert out.numel() == x.shape[1]
    _call(dev, "fc_colsum_bf16", ptr(x), x.stride(0), x.shape[0], x.shape[1], ptr(out))
    return out


def transpose(x: torch.Tensor, group_len: int = 0, group_skip: int = 0, colsum: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 ``(rows, cols)`` -> ``(cols, pad8(kept rows))`` with zero-filled padding columns; ``colsum += x.sum(0)``."""
    dev = _dev(x)
    assert x.dtype == BF16 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    kept = rows if group_len == 0 else rows // group_len * (group_len - group_skip)
    ld = pad8(kept)
    if out is None:
        out = torch.empty(cols, ld, device=dev, dtype=BF16)
    assert out.shape == (cols, ld) and out.is_contiguous()
    _call(dev, "fc_transpose_bf16", ptr(x), x.stride(0), ptr(out), ld, kept, cols, group_len, group_skip, ptr(colsum))
    return out


def layernorm_bwd(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor,
                  add: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                  eps: float = 1e-5) -> torch.Tensor:
    dev = _dev(x)
    assert x.dtype == BF16 and dy.dtype == BF16 and x.is_contiguous() and dy.is_contiguous() and x.shape == dy.shape
    assert gamma.dtype == torch.float32 and dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32
    assert add is None or (add.dtype == BF16 and add.is_contiguous() and add.shape == x.shape)
    out = torch.empty_like(x) if out is None else out
    rows, D = x.shape
    _call(dev, "fc_layernorm_bwd_bf16", ptr(x), ptr(dy), ptr(gamma), ptr(add), ptr(out), ptr(dgamma), ptr(dbeta),
          rows, D, eps)
    return out


def quickgelu(u: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _dev(u)
    assert u.dtype == BF16 and u.is_contiguous()
    out = torch.empty_like(u) if out is None else out
    _call(dev, "fc_quickgelu_bf16", ptr(u), ptr(out), u.numel())
    return out


def quickgelu_bwd(u: torch.Tensor, dg: torch.Tensor, out: Optional[torch.Tensor] = None,
                  g_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``du = dg * quickgelu'(u)``; ``g_out`` (optional) also receives ``quickgelu(u)`` from the same pass."""
    dev = _dev(u)
    assert u.dtype == BF16 and dg.dtype == BF16 and u.is_contiguous() and dg.is_contiguous() and u.shape == dg.shape
    assert g_out is None or (g_out.dtype == BF16 and g_out.is_contiguous() and g_out.shape == u.shape)
    out = torch.empty_like(u) if out is None else out
    _call(dev, "fc_quickgelu_bwd_bf16", ptr(u), ptr(dg), ptr(out), ptr(g_out), u.numel())
    return out


def attention_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, seqs: int, L: int, heads: int,
                  causal: bool, dqkv: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _dev(qkv)
    assert qkv.dtype == BF16 and qkv.is_contiguous() and qkv.shape == (seqs * L, 3 * heads * 64)
    assert out.dtype == BF16 and out.is_contiguous() and out.shape == (seqs * L, heads * 64)
    assert dout.dtype == BF16 and dout.is_contiguous() and dout.shape == out.shape
    dqkv = torch.empty_like(qkv) if dqkv is None else dqkv
    _call(dev, "fc_attention_bwd_bf16", ptr(qkv), ptr(out), ptr(dout), ptr(dqkv), seqs, L, heads, int(causal))
    return dqkv


def loss_fwd_bwd(scores: torch.Tensor, teacher_scores: Optional[torch.Tensor] = None, gscale: float = 1.0,
                 want_grad: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """-> (loss (), ``gscale * dloss/dscores`` or None).  ``teacher_scores`` None: ``nce_loss``; else the
    teacher-student KL form with ``batchmean`` reduction (``aligner/loss.py:13-39``), which also takes rectangular
    (videos x prompts) matrices."""
    dev = _dev(scores)
    R, C = scores.shape
    assert scores.dtype == torch.float32 and scores.stride(1) == 1
    if teacher_scores is not None:
        assert teacher_scores.dtype == torch.float32 and teacher_scores.shape == (R, C)
        assert teacher_scores.stride() == scores.stride()
    else:
        assert R == C, "nce_loss pairs row i with column i"
    lse = torch.empty(2 * (R + C), device=dev, dtype=torch.float32)
    loss = torch.empty((), device=dev, dtype=torch.float32)
    grad = torch.empty(R, C, device=dev, dtype=torch.float32) if want_grad else None
    _call(dev, "fc_loss_fwd_bwd", ptr(scores), ptr(teacher_scores), scores.stride(0), R, C, ptr(lse), gscale, ptr(loss),
          ptr(grad), C)
    return loss, grad


def sgemm(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
          alpha: float = 1.0) -> torch.Tensor:
    """``alpha * op(a) @ op(b)`` in fp32, ``op(x) = x.T`` when ``trans_x``."""
    dev = _dev(a)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    assert (b.shape[1] if trans_b else b.shape[0]) == K
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    # the kernel's trans_b flag means "B is stored (N, K)"
    _call(dev, "fc_sgemm_f32", int(trans_a), int(trans_b), M, N, K, alpha, ptr(a), a.stride(0), ptr(b), b.stride(0),
          ptr(out), N)
    return out


def pool_normalize_bwd(x: torch.Tensor, dout: torch.Tensor, frames_per_row: int, scale: float = 1.0) -> torch.Tensor:
    dev = _dev(x)
    assert x.dtype == torch.float32 and dout.dtype == torch.float32 and x.is_contiguous() and dout.is_contiguous()
    B, D = dout.shape
    assert x.shape == (B * frames_per_row, D)
    dx = torch.empty(x.shape, device=dev, dtype=BF16)
    _call(dev, "fc_pool_normalize_bwd", ptr(x), ptr(dout), ptr(dx), B, frames_per_row, D, scale)
    return dx


def gather_seq_rows(x: torch.Tensor, ids: Optional[torch.Tensor], seqs: int, L: int) -> torch.Tensor:
    dev = _dev(x)
    W = x.shape[1]
    assert x.dtype == BF16 and x.is_contiguous() and x.shape[0] == seqs * L
    rows = torch.empty(seqs, W, device=dev, dtype=BF16)
    _call(dev, "fc_seq_rows", ptr(x), ptr(ids), ptr(rows), seqs, L, W, 0)
    return rows


def scatter_seq_rows(rows: torch.Tensor, ids: Optional[torch.Tensor], L: int) -> torch.Tensor:
    dev = _dev(rows)
    seqs, W = rows.shape
    assert rows.dtype == BF16 and rows.is_contiguous()
    x = torch.zeros(seqs * L, W, device=dev, dtype=BF16)
    _call(dev, "fc_seq_rows", ptr(x), ptr(ids), ptr(rows), seqs, L, W, 1)
    return x


def seq_sum(dx: torch.Tensor, out: torch.Tensor, seqs: int, L: int) -> torch.Tensor:
    dev = _dev(dx)
    W = dx.shape[1]
    assert dx.dtype == BF16 and dx.is_contiguous() and dx.shape[0] == seqs * L
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == L * W
    _call(dev, "fc_seq_sum", ptr(dx), ptr(out), seqs, L, W)
    return out


def token_scatter_add(ids: torch.Tensor, dx: torch.Tensor, dtok: torch.Tensor) -> torch.Tensor:
    dev = _dev(dx)
    assert ids.dtype == torch.int32 and ids.is_contiguous() and dx.dtype == BF16 and dx.is_contiguous()
    assert dtok.dtype == torch.float32 and dtok.is_contiguous() and dtok.shape[1] == dx.shape[1]
    _call(dev, "fc_token_scatter_add", ptr(ids), ptr(dx), ptr(dtok), ids.numel(), dx.shape[1], dtok.shape[0])
    return dtok


def adamw_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
               betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
               p_bf16: Optional[torch.Tensor] = None) -> None:
    dev = _dev(p)
    for t in (p, g, m, v):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == p.numel()
    _call(dev, "fc_adamw_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), p.numel(), lr, betas[0], betas[1], eps,
          weight_decay, step)


def f32_to_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _dev(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(x.shape, device=dev, dtype=BF16) if out is None else out
    _call(dev, "fc_f32_to_bf16", ptr(x), ptr(out), x.numel())
    return out


def patch_embed(frames: torch.Tensor, conv_w: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor,
                patch: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """frames ``(F,3,R,R)`` fp32/bf16 -> (x bf16 ``(F*(G*G+1), W)``, patches bf16 ``(F*G*G, 3*P*P)``)."""
    dev = _dev(frames)
    assert frames.is_contiguous() and frames.dtype in _lib.DTYPE_CODE and conv_w.dtype == BF16
    F, _, R, _ = frames.shape
    W = conv_w.shape[0]
    G = R // patch
    patches = torch.empty(F * G * G, 3 * patch * patch, device=dev, dtype=BF16)
    x = torch.empty(F * (G * G + 1), W, device=dev, dtype=BF16)
    _call(dev, "fc_patch_embed", ptr(frames), _lib.DTYPE_CODE[frames.dtype], ptr(conv_w), ptr(cls), ptr(pos),
          ptr(patches), ptr(x), F, R, patch, W)
    return x, patches


def text_embed(ids: torch.Tensor, tok: torch.Tensor, pos: torch.Tensor, err_flag: torch.Tensor) -> torch.Tensor:
    dev = _dev(ids)
    assert ids.dtype == torch.int32 and ids.is_contiguous() and tok.dtype == torch.float32
    C, L = ids.shape
    W = tok.shape[1]
    x = torch.empty(C * L, W, device=dev, dtype=BF16)
    _call(dev, "fc_text_embed", ptr(ids), ptr(tok), ptr(pos), ptr(x), C, L, W, tok.shape[0], ptr(err_flag))
    return x
