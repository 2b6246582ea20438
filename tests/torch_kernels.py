"""TEST INFRASTRUCTURE ONLY: torch stand-ins for every kernel the trainer (``fitclip_b200/training.py``) chains, with the
same signatures as ``fitclip_b200.training.NativeKernels``, in fp32 on whatever device the tensors live on.

They let the ORCHESTRATION of the explicit backward pass (which gradient goes where, transposes, padding, residual
adds, head / embedding plumbing, gathers across ranks) be checked against torch.autograd on CPU, where no CUDA kernel
can run.  Each stand-in is the textbook formula of its kernel, NOT autograd, so a wrong formula in the plan shows up
too.  Nothing under ``fitclip_b200/`` imports this module; the product path has no fallback."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class TorchKernels:
    act_dtype = torch.float32

    @staticmethod
    def cast(x, out=None):
        if out is None:
            return x.clone()
        out.copy_(x)
        return out

    @staticmethod
    def linear(a, w, bias=None, resid=None):
        y = a @ w.T
        if bias is not None:
            y = y + bias
        return y if resid is None else y + resid

    @staticmethod
    def linear_nt(a, w, resid=None):
        return a @ w if resid is None else a @ w + resid

    @staticmethod
    def matmul_nt_f32(a, w):
        return a @ w

    @staticmethod
    def linear_nt_qgelu_bwd(a, w, u):
        s = torch.sigmoid(1.702 * u)
        return (a @ w) * (s * (1 + 1.702 * u * (1 - s)))

    @staticmethod
    def wgrad_tn(dy, x, out, alpha=1.0, k_splits=0, rows=None):
        rows = dy.shape[0] if rows is None else rows
        out += alpha * (dy[:rows].T @ x[:rows])
        return out

    @staticmethod
    def colsum(x, out):
        out += x.sum(0)
        return out

    @staticmethod
    def layernorm(x, g, b, eps=1e-5, out=None):
        mean = x.mean(-1, keepdim=True)
        var = ((x - mean) ** 2).mean(-1, keepdim=True)
        return (x - mean) * torch.rsqrt(var + eps) * g + b

    @staticmethod
    def layernorm_bwd(x, dy, gamma, dgamma, dbeta, add=None, out=None, eps=1e-5):
        mean = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(((x - mean) ** 2).mean(-1, keepdim=True) + eps)
        xh = (x - mean) * rstd
        g = dy * gamma
        dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        dgamma += (dy * xh).sum(0)
        dbeta += dy.sum(0)
        dx = dx if add is None else dx + add
        if out is not None:
            out.copy_(dx)
            return out
        return dx

    @staticmethod
    def _split(qkv, seqs, L, heads):
        return qkv.view(seqs, L, 3, heads, 64).permute(2, 0, 3, 1, 4)  # (3, S, H, L, 64)

    @classmethod
    def _probs(cls, q, k, L, causal):
        s = (q @ k.transpose(-1, -2)) * 0.125
        if causal:
            s = s + torch.full((L, L), float("-inf"), device=s.device).triu_(1)
        return torch.softmax(s, dim=-1)

    @classmethod
    def attention(cls, qkv, seqs, L, heads, causal):
        q, k, v = cls._split(qkv, seqs, L, heads)
        o = cls._probs(q, k, L, causal) @ v
        return o.permute(0, 2, 1, 3).reshape(seqs * L, heads * 64)

    @classmethod
    def attention_bwd(cls, qkv, out, dout, seqs, L, heads, causal, dqkv=None):
        q, k, v = cls._split(qkv, seqs, L, heads)
        p = cls._probs(q, k, L, causal)
        do = dout.view(seqs, L, heads, 64).permute(0, 2, 1, 3)
        o = out.view(seqs, L, heads, 64).permute(0, 2, 1, 3)
        dv = p.transpose(-1, -2) @ do
        dp = do @ v.transpose(-1, -2)
        delta = (do * o).sum(-1, keepdim=True)
        ds = p * (dp - delta)
        dq = ds @ k * 0.125
        dk = ds.transpose(-1, -2) @ q * 0.125
        return torch.stack([dq, dk, dv]).permute(1, 3, 0, 2, 4).reshape(seqs * L, 3 * heads * 64)

    @staticmethod
    def quickgelu(u, out=None):
        return u * torch.sigmoid(1.702 * u)

    empty_like = staticmethod(torch.empty_like)

    @staticmethod
    def quickgelu_bwd(u, dg, out=None, g_out=None):
        s = torch.sigmoid(1.702 * u)
        du = dg * s * (1 + 1.702 * u * (1 - s))
        if g_out is not None:
            g_out.copy_(u * s)
        if out is not None:
            out.copy_(du)
            return out
        return du

    @staticmethod
    def transpose(x, group_len=0, group_skip=0, colsum=None, out=None):
        rows, cols = x.shape
        kept = x if group_len == 0 else x.view(rows // group_len, group_len, cols)[:, group_skip:].reshape(-1, cols)
        if colsum is not None:
            colsum += kept.sum(0)
        res = torch.zeros(cols, pad8(kept.shape[0]), dtype=x.dtype, device=x.device)
        res[:, :kept.shape[0]] = kept.T
        if out is not None:
            out.copy_(res)
            return out
        return res

    @staticmethod
    def wgrad(a, b, out, alpha=1.0, k=None, k_splits=0):
        k = a.shape[1] if k is None else k
        out += alpha * (a[:, :k] @ b[:, :k].T)
        return out

    @staticmethod
    def patch_embed(frames, conv_w, cls, pos, patch):
        Fr, _, R, _ = frames.shape
        W = conv_w.shape[0]
        G = R // patch
        patches = frames.view(Fr, 3, G, patch, G, patch).permute(0, 2, 4, 1, 3, 5).reshape(Fr * G * G, 3 * patch * patch)
        tok = (patches @ conv_w.T).view(Fr, G * G, W)
        x = torch.cat([cls.expand(Fr, 1, W), tok], 1) + pos
        return x.reshape(Fr * (G * G + 1), W), patches

    @staticmethod
    def text_embed(ids, tok, pos):
        return (tok[ids.long()] + pos).reshape(-1, tok.shape[1])

    @staticmethod
    def _eot(ids, seqs):
        return torch.zeros(seqs, dtype=torch.long) if ids is None else ids.argmax(-1)

    @classmethod
    def gather_seq_rows(cls, x, ids, seqs, L):
        return x.view(seqs, L, -1)[torch.arange(seqs), cls._eot(ids, seqs)].clone()

    @classmethod
    def scatter_seq_rows(cls, rows, ids, L):
        seqs, W = rows.shape
        x = torch.zeros(seqs, L, W, dtype=rows.dtype)
        x[torch.arange(seqs), cls._eot(ids, seqs)] = rows
        return x.view(seqs * L, W)

    @staticmethod
    def seq_sum(dx, out, seqs, L):
        out += dx.view(seqs, L, -1).sum(0).view(out.shape)
        return out

    @staticmethod
    def token_scatter_add(ids, dx, dtok):
        dtok.index_add_(0, ids.long(), dx)
        return dtok

    @staticmethod
    def pool_normalize(x, T, scale=1.0):
        return scale * (x / x.norm(dim=-1, keepdim=True)).view(-1, T, x.shape[1]).mean(1)

    @staticmethod
    def pool_normalize_bwd(x, dout, T, scale=1.0):
        n = x.norm(dim=-1, keepdim=True)
        xh = x / n
        d = dout.repeat_interleave(T, 0)
        return scale / T * (d - xh * (xh * d).sum(-1, keepdim=True)) / n

    @staticmethod
    def sgemm(a, b, trans_a=False, trans_b=False, alpha=1.0):
        return alpha * ((a.T if trans_a else a) @ (b.T if trans_b else b))

    @staticmethod
    def loss_fwd_bwd(scores, teacher_scores=None, gscale=1.0, want_grad=True):
        B, C = scores.shape  # rows = videos, columns = texts (C != B only with prompts, teacher-student form)
        pr, pc = torch.softmax(scores, 1), torch.softmax(scores, 0)
        if teacher_scores is None:
            loss = (-F.log_softmax(scores, 1).diag()).mean() + (-F.log_softmax(scores, 0).diag()).mean()
            grad = (pr + pc - 2 * torch.eye(B)) / B
        else:
            tr, tc = torch.softmax(teacher_scores, 1), torch.softmax(teacher_scores, 0)
            loss = (tr * (F.log_softmax(teacher_scores, 1) - F.log_softmax(scores, 1))).sum() / B \
                + (tc * (F.log_softmax(teacher_scores, 0) - F.log_softmax(scores, 0))).sum() / C
            grad = (pr - tr) / B + (pc - tc) / C
        return loss, gscale * grad

    @staticmethod
    def adamw_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, p_bf16=None):
        b1, b2 = betas
        p.mul_(1 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
        if p_bf16 is not None:
            p_bf16.copy_(p)
