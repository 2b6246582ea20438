"""Teacher-student training step on B200.

Reference: ``TeacherStudentLightningModule`` (``aligner/teacher_student.py:43-183``): student and frozen teacher encode
the same batch (``_step``, ``:93-96``); per dataset the scaled ``B x B`` score matrix feeds ``nce_loss`` (labelled) or
``TeacherStudentNCELoss("batchmean") * exp(ts_scale)^2`` (unlabelled) (``_dataset_step_end``, ``:142-173``); the
dataset losses are mixed with ``dataset_loss_share`` (``training_step_end``, ``:176-183``); Lightning then runs
``loss.backward()`` and ``torch.optim.AdamW(lr=3e-6).step()`` (``config/trainer.yaml:22-24``, ``aligner/cli.py:126-134``).

The reference gets the backward pass from torch.autograd.  Here it is an explicit chain of the kernels in
``csrc/train.cu`` plus the tcgen05 GEMM (``fitclip_b200.train_ops``):

* all parameters of the student live in ONE flat fp32 buffer (the ``nn.Parameter``s of the :class:`B200Clip` become
  views into it), with flat fp32 gradient / Adam-moment buffers and a flat bf16 mirror next to it: the optimizer is one
  launch, the gradient all-reduce across GPUs one NCCL call, and ``zero_grad`` one memset;
* the forward keeps the activations the backward needs (per block: block input, QKV, attention output, post-attention
  residual, MLP pre-activation and both LayerNorm outputs = 12 x tokens x width bf16; 89 GB for 2048 frames of
  ViT-B/16, which is what the 180 GB of HBM are for) -- nothing is recomputed except QuickGELU values, which the
  backward kernel emits on its single pass over the pre-activation (``keep_layernorm=False`` drops the LayerNorm
  outputs and recomputes them: 15 GB less, 2 % slower);
* the backward GEMMs read their operands IN PLACE through MN-major tcgen05 descriptors: dgrad ``dX = dY W`` takes ``W``
  as the forward pass stores it, wgrad ``dW = dY^T X`` takes ``dY`` and ``X`` as they are and splits K (= the token
  count) across the SMs -- no transposed copies of weights or activations exist (only the patch-embedding weight
  gradient, once per step, goes through a transpose because its rows skip the class token).

The kernel set is injected (``kernels=``) so that the orchestration can be checked on CPU against torch.autograd with
torch stand-ins for every kernel (``tests/torch_kernels.py``, test infrastructure); the default and only product path
is :class:`NativeKernels` -- there is no fallback.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import torch

from . import _lib, ops, train_ops as T
from .encoder import B200Clip


class NativeKernels:
    """The product kernel set: every call lands in ``libfitclip_b200.so``."""
    act_dtype = torch.bfloat16

    def __init__(self, device: torch.device) -> None:
        self.device = device
        self._zeros: Dict[int, torch.Tensor] = {}
        self.err_flag = torch.zeros(64, device=device, dtype=torch.int32)

    def zero_bias(self, n: int) -> torch.Tensor:
        if n not in self._zeros:
            self._zeros[n] = torch.zeros(n, device=self.device, dtype=torch.float32)
        return self._zeros[n]

    cast = staticmethod(T.f32_to_bf16)
    empty_like = staticmethod(torch.empty_like)

    def linear(self, a, w, bias=None, resid=None):
        bias = self.zero_bias(w.shape[0]) if bias is None else bias
        return ops.gemm_bf16(a, w, bias, resid, _lib.EPI_BIAS if resid is None else _lib.EPI_BIAS_RESID)

    def linear_nt(self, a, w, resid=None):
        """``a @ w`` (+ resid) with ``w (K, N)`` read in place (MN-major B operand)."""
        return T.gemm_nt(a, w, self.zero_bias(w.shape[1]), resid)

    def matmul_nt_f32(self, a, w):
        return T.gemm_nt(a, w, None, f32=True)

    def linear_nt_qgelu_bwd(self, a, w, u):
        """``(a @ w) * quickgelu'(u)``: c_proj's dgrad with QuickGELU's backward in its epilogue."""
        return T.gemm_nt(a, w, None, qgelu_bwd_of=u)

    wgrad_tn = staticmethod(T.wgrad_tn)
    colsum = staticmethod(T.colsum)

    layernorm = staticmethod(ops.layernorm_bf16)
    layernorm_bwd = staticmethod(T.layernorm_bwd)
    attention = staticmethod(ops.attention_bf16)
    attention_bwd = staticmethod(T.attention_bwd)
    quickgelu = staticmethod(T.quickgelu)
    quickgelu_bwd = staticmethod(T.quickgelu_bwd)
    transpose = staticmethod(T.transpose)
    wgrad = staticmethod(T.gemm_splitk)
    patch_embed = staticmethod(T.patch_embed)
    gather_seq_rows = staticmethod(T.gather_seq_rows)
    scatter_seq_rows = staticmethod(T.scatter_seq_rows)
    seq_sum = staticmethod(T.seq_sum)
    token_scatter_add = staticmethod(T.token_scatter_add)
    pool_normalize = staticmethod(ops.pool_normalize)
    pool_normalize_bwd = staticmethod(T.pool_normalize_bwd)
    sgemm = staticmethod(T.sgemm)
    loss_fwd_bwd = staticmethod(T.loss_fwd_bwd)
    adamw_step = staticmethod(T.adamw_step)

    def text_embed(self, ids, tok, pos):
        return T.text_embed(ids, tok, pos, self.err_flag)


class ClipTrainer:
    """Forward-with-saved-activations, backward and AdamW for one :class:`B200Clip` (the student)."""

    def __init__(self, model: B200Clip, lr: float = 3e-6, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, kernels: Any = None, keep_layernorm: bool = True,
                 keep_activation: bool = True, gradient_clip_val: Optional[float] = None) -> None:
        self.model = model
        # Lightning's Trainer(gradient_clip_val=...) (config/trainer.yaml:39, null in the reference's config): the global
        # L2 norm of ALL gradients of the optimizer is brought down to this value before the step (clip_grad_norm_)
        self.gradient_clip_val = gradient_clip_val
        self.keep_layernorm = keep_layernorm
        # keep QuickGELU(u) from the forward (+ 2 B per MLP-hidden element: 30 GB at 2048 frames of ViT-B/16, 133 GB peak)
        # instead of re-emitting it from the backward's pass over u: that pass then writes 2 B per element less
        self.keep_activation = keep_activation
        self.cfg = dict(model.config)
        if self.cfg.get("vision_tower", 0) != 0:
            raise NotImplementedError("the training step is built for OpenAI-layout students (ln_pre, QuickGELU); a "
                                      "SLIP-layout model (timm tower) can be the frozen TEACHER, not the student")
        if (3 * self.cfg["vision_patch_size"] ** 2) % 8:
            raise ValueError("training needs 3 * patch_size^2 to be a multiple of 8 (ViT-B/16, ViT-B/32)")
        named = [(n, p) for n, p in model.named_parameters() if n != "logit_scale"]
        device = named[0][1].device
        self.K = NativeKernels(device) if kernels is None else kernels
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        # ---- one flat fp32 master buffer; the module's parameters become views into it
        total = sum((p.numel() + 63) // 64 * 64 for _, p in named)  # 256-byte aligned slots (TMA / 16-byte loads)
        self.flat = torch.zeros(total, device=device, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.flat_act = torch.zeros(total, device=device, dtype=self.K.act_dtype)
        self.w: Dict[str, torch.Tensor] = {}   # fp32 parameter views
        self.g: Dict[str, torch.Tensor] = {}   # fp32 gradient views
        self.wb: Dict[str, torch.Tensor] = {}  # act-dtype (bf16) views of the mirror, GEMM weights as (N, K)
        off = 0
        for name, p in named:
            n = p.numel()
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view(p.shape)
            p.requires_grad_(False)
            self.w[name] = p.data
            self.g[name] = self.grad[off:off + n].view(p.shape)
            self.wb[name] = self.flat_act[off:off + n].view(p.shape)
            off += (n + 63) // 64 * 64
        self.refresh_weight_copies()

    # ------------------------------------------------------------------------------------------------ weights
    def refresh_weight_copies(self) -> None:
        """bf16 mirror of the flat fp32 parameters (needed after an external edit of the parameters; AdamW refreshes
        it itself)."""
        self.K.cast(self.flat, out=self.flat_act)

    def zero_grad(self) -> None:
        self.grad.zero_()

    def check_inputs(self) -> None:
        """Synchronises and raises if a token id fed to the training forward was out of range: ``text_embed_kernel``
        clamps such an id to token 0 and raises the flag, ``token_scatter_kernel`` drops its gradient -- torch's embedding
        lookup in the reference would have raised.  Resets the flag."""
        flag = getattr(self.K, "err_flag", None)
        if flag is not None and bool(flag.any().item()):
            flag.zero_()
            raise _lib.FitclipError(-3, "training step: token id out of range (text_embed clamped it to 0)")

    def optimizer_step(self, group=None, extra_grads: Sequence[torch.Tensor] = ()) -> None:
        """All-reduce(SUM) of the flat gradient across ``group`` (each rank holds the gradient of the GLOBAL-batch loss
        through its own samples), optional global-norm clipping (``extra_grads``: gradients of parameters that live outside
        the flat buffer -- the logit scales -- which share the norm and the coefficient), then one fused AdamW launch and
        the weight-copy refresh."""
        dist = torch.distributed
        if group is not False and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, group=group)
        if self.gradient_clip_val is not None:  # torch.nn.utils.clip_grad_norm_(params, max_norm), norm type 2
            sq = self.grad.double().pow(2).sum()
            for g in extra_grads:
                sq = sq + g.double().pow(2).sum()
            coef = (self.gradient_clip_val / (sq.sqrt() + 1e-6)).clamp(max=1.0).float()
            self.grad.mul_(coef)
            for g in extra_grads:
                g.mul_(coef)
        self.step_count += 1
        self.K.adamw_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas,
                          self.eps, self.weight_decay, p_bf16=self.flat_act)
        # the evaluation engine of the module re-uploads its folded copies when it next runs
        self.model._engine.signature = None

    # ------------------------------------------------------------------------------------------------ blocks
    def _blocks_forward(self, x, prefix: str, layers: int, seqs: int, L: int, heads: int, causal: bool, saved: List):
        K, w, wb = self.K, self.w, self.wb
        for i in range(layers):
            p = f"{prefix}resblocks.{i}."
            ln1 = K.layernorm(x, w[p + "ln_1.weight"], w[p + "ln_1.bias"])
            qkv = K.linear(ln1, wb[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"])
            att = K.attention(qkv, seqs, L, heads, causal)
            x_mid = K.linear(att, wb[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"], resid=x)
            ln2 = K.layernorm(x_mid, w[p + "ln_2.weight"], w[p + "ln_2.bias"])
            u = K.linear(ln2, wb[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"])
            act = K.quickgelu(u)
            x_out = K.linear(act, wb[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"], resid=x_mid)
            saved.append((x, qkv, att, x_mid, u) + ((ln1, ln2) if self.keep_layernorm else (None, None))
                         + ((act,) if self.keep_activation else (None,)))
            x = x_out
        return x

    def _linear_backward(self, dy, x_in, name_w: str, name_b: str, want_dx: bool = True):
        """Gradients of ``y = x_in @ W.T + b``: db += colsum(dy), dW += dy^T x_in, returns dx = dy @ W."""
        K = self.K
        K.colsum(dy, self.g[name_b])
        K.wgrad_tn(dy, x_in, self.g[name_w])
        return K.linear_nt(dy, self.wb[name_w]) if want_dx else None

    def _blocks_backward(self, dx, prefix: str, layers: int, seqs: int, L: int, heads: int, causal: bool, saved: List):
        K, w, g = self.K, self.w, self.g
        for i in reversed(range(layers)):
            p = f"{prefix}resblocks.{i}."
            x_in, qkv, att, x_mid, u, ln1, ln2, act = saved.pop()
            # x_out = x_mid + c_proj(quickgelu(c_fc(ln_2(x_mid))))
            # dact first (it does not need quickgelu(u)); then ONE pass over u gives du -- and, unless the forward kept it,
            # quickgelu(u), which the c_proj weight gradient reads
            if act is None:
                dact = K.linear_nt(dx, self.wb[p + "mlp.c_proj.weight"])
                act = K.empty_like(u)
                du = K.quickgelu_bwd(u, dact, out=dact, g_out=act)
            else:  # the forward kept quickgelu(u): du = (dx W) o quickgelu'(u) leaves the dgrad GEMM's epilogue directly
                dact = du = K.linear_nt_qgelu_bwd(dx, self.wb[p + "mlp.c_proj.weight"], u)
            K.colsum(dx, g[p + "mlp.c_proj.bias"])
            K.wgrad_tn(dx, act, g[p + "mlp.c_proj.weight"])
            del act
            if ln2 is None:
                ln2 = K.layernorm(x_mid, w[p + "ln_2.weight"], w[p + "ln_2.bias"])
            dln2 = self._linear_backward(du, ln2, p + "mlp.c_fc.weight", p + "mlp.c_fc.bias")
            del du, dact, ln2
            dx_mid = K.layernorm_bwd(x_mid, dln2, w[p + "ln_2.weight"], g[p + "ln_2.weight"], g[p + "ln_2.bias"],
                                     add=dx, out=dx)
            # x_mid = x_in + out_proj(attention(in_proj(ln_1(x_in))))
            datt = self._linear_backward(dx_mid, att, p + "attn.out_proj.weight", p + "attn.out_proj.bias")
            dqkv = K.attention_bwd(qkv, att, datt, seqs, L, heads, causal)
            if ln1 is None:
                ln1 = K.layernorm(x_in, w[p + "ln_1.weight"], w[p + "ln_1.bias"])
            dln1 = self._linear_backward(dqkv, ln1, p + "attn.in_proj_weight", p + "attn.in_proj_bias")
            del dqkv, datt, ln1
            dx = K.layernorm_bwd(x_in, dln1, w[p + "ln_1.weight"], g[p + "ln_1.weight"], g[p + "ln_1.bias"],
                                 add=dx_mid, out=dx_mid)
        return dx

    # ------------------------------------------------------------------------------------------------ heads
    def _head_forward(self, x, ids, seqs: int, L: int, ln: str, proj: str, pool: int, ctx: Dict) -> torch.Tensor:
        K, w = self.K, self.w
        rows = K.gather_seq_rows(x, ids, seqs, L)
        normed = K.layernorm(rows, w[ln + ".weight"], w[ln + ".bias"])
        feat = K.matmul_nt_f32(normed, self.wb[proj])  # (seqs, E) fp32 = ln(x[eot]) @ proj
        ctx.update(rows=rows, normed=normed, feat=feat, ids=ids, seqs=seqs, L=L, pool=pool)
        return K.pool_normalize(feat, pool)

    def _head_backward(self, dpooled, ln: str, proj: str, ctx: Dict):
        K, w, g = self.K, self.w, self.g
        dfeat = K.pool_normalize_bwd(ctx["feat"], dpooled.contiguous(), ctx["pool"])
        # feat = normed @ proj with proj (W, E):  dproj += normed^T dfeat,  dnormed = dfeat @ proj^T
        K.wgrad_tn(ctx["normed"], dfeat, g[proj])
        dnormed = K.linear(dfeat, self.wb[proj])
        drows = K.layernorm_bwd(ctx["rows"], dnormed, w[ln + ".weight"], g[ln + ".weight"], g[ln + ".bias"])
        return K.scatter_seq_rows(drows, ctx["ids"], ctx["L"])

    # ------------------------------------------------------------------------------------------------ towers
    def encode_video(self, video: torch.Tensor) -> torch.Tensor:
        """``ClipVideoTextEncoder.encode_video`` (clip_video_text_encoder.py:80-89) keeping what backward needs."""
        K, w, c = self.K, self.w, self.cfg
        B, Tn = video.shape[:2]
        F = B * Tn
        G = c["image_resolution"] // c["vision_patch_size"]
        L, W = G * G + 1, c["vision_width"]
        frames = video.reshape(F, *video.shape[2:]).contiguous()
        conv = self.wb["visual.conv1.weight"].view(W, -1)
        x0, patches = K.patch_embed(frames, conv, w["visual.class_embedding"], w["visual.positional_embedding"],
                                    c["vision_patch_size"])
        x = K.layernorm(x0, w["visual.ln_pre.weight"], w["visual.ln_pre.bias"])
        saved: List = []
        x = self._blocks_forward(x, "visual.transformer.", c["vision_layers"], F, L, W // 64, False, saved)
        head: Dict = {}
        out = self._head_forward(x, None, F, L, "visual.ln_post", "visual.proj", Tn, head)
        self._vis = dict(x0=x0, patches=patches, saved=saved, head=head, F=F, L=L, W=W)
        return out

    def backward_video(self, dvideo: torch.Tensor) -> None:
        K, w, g, c, v = self.K, self.w, self.g, self.cfg, self._vis
        F, L, W = v["F"], v["L"], v["W"]
        dx = self._head_backward(dvideo, "visual.ln_post", "visual.proj", v["head"])
        dx = self._blocks_backward(dx, "visual.transformer.", c["vision_layers"], F, L, W // 64, False, v["saved"])
        dx0 = K.layernorm_bwd(v["x0"], dx, w["visual.ln_pre.weight"], g["visual.ln_pre.weight"],
                              g["visual.ln_pre.bias"], out=dx)
        # x0[f, 0] = class_embedding + pos[0];  x0[f, 1 + p] = conv1(patch p) + pos[1 + p]
        K.seq_sum(dx0, g["visual.positional_embedding"], F, L)
        K.seq_sum(K.gather_seq_rows(dx0, None, F, L), g["visual.class_embedding"], F, 1)
        gconv = g["visual.conv1.weight"]
        K.wgrad(K.transpose(dx0, group_len=L, group_skip=1), K.transpose(v["patches"]), gconv.view(W, -1))
        self._vis = None

    def encode_text(self, ids: torch.Tensor) -> torch.Tensor:
        """``ClipVideoTextEncoder.encode_text`` (clip_video_text_encoder.py:92-94) keeping what backward needs."""
        K, w, c = self.K, self.w, self.cfg
        ids = ids.to(torch.int32).contiguous()
        C, L = ids.shape
        W = c["transformer_width"]
        x = K.text_embed(ids, w["token_embedding.weight"], w["positional_embedding"])
        saved: List = []
        x = self._blocks_forward(x, "transformer.", c["transformer_layers"], C, L, c["transformer_heads"], True, saved)
        head: Dict = {}
        out = self._head_forward(x, ids, C, L, "ln_final", "text_projection", 1, head)
        self._txt = dict(ids=ids, saved=saved, head=head, C=C, L=L)
        return out

    def backward_text(self, dtext: torch.Tensor) -> None:
        K, g, c, t = self.K, self.g, self.cfg, self._txt
        C, L = t["C"], t["L"]
        dx = self._head_backward(dtext, "ln_final", "text_projection", t["head"])
        dx = self._blocks_backward(dx, "transformer.", c["transformer_layers"], C, L, c["transformer_heads"], True,
                                   t["saved"])
        K.seq_sum(dx, g["positional_embedding"], C, L)
        K.token_scatter_add(t["ids"].view(-1), dx, g["token_embedding.weight"])
        self._txt = None


def _all_gather_rows(t: torch.Tensor, group) -> Tuple[torch.Tensor, int]:
    """(world * B, D) concatenation of every rank's (B, D) rows + this rank's row offset (``all_gather`` of
    ``util/tensor_utils.py:48-66``; equal local batch sizes, as the reference's DistributedSampler gives)."""
    dist = torch.distributed
    if group is False or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t, 0
    out = torch.empty(dist.get_world_size(group) * t.shape[0], *t.shape[1:], device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out, dist.get_rank(group) * t.shape[0]


def _adamw_scales_step(tr: "ClipTrainer", temps: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
                       max_logit_scale: float) -> None:
    """AdamW on the log-space logit scale(s) with the trainer's hyper-parameters (one parameter group in the reference,
    ``aligner/cli.py:126-134``), then the clamp of ``optimizer_step`` (video_text_module.py:93-97)."""
    b1, b2 = tr.betas
    step = tr.step_count  # already advanced by ClipTrainer.optimizer_step
    temps.mul_(1 - tr.lr * tr.weight_decay)
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + tr.eps
    temps.addcdiv_(m, denom, value=-tr.lr / (1 - b1 ** step))
    temps.clamp_(max=max_logit_scale)


class VideoTextTrainingModule:
    """``VideoTextLightningModule`` training path without Lightning (``aligner/video_text_module.py:25-97``; what
    ``command=train`` runs for a single encoder with the default ``TextVideoRetrievalLightningModule``): one encoder, the
    embeddings of the step gathered across ranks, ``scores = exp(logit_scale) * V @ T.T``, ``NCELoss`` (``loss.py:13-26``),
    backward, AdamW; the logit scale is a trained parameter by default (``fit_temperature=True``, ``:26-34``) and is clamped
    at ``-log(min_temperature)`` after every optimizer step (``:93-97``).

    ``batch``: ``video (B,T,3,R,R)``, ``text {"input_ids": (B, 77)}`` (``video_id`` is ignored, ``:41``)."""

    def __init__(self, encoder, init_temperature: float = 0.05, min_temperature: float = 0.001,
                 fit_temperature: bool = True, lr: float = 3e-6, weight_decay: float = 1e-2, group=None,
                 kernels: Any = None, gradient_clip_val: Optional[float] = None) -> None:
        self.encoder = encoder
        self.trainer = ClipTrainer(encoder.model, lr=lr, weight_decay=weight_decay, kernels=kernels,
                                   gradient_clip_val=gradient_clip_val)
        self.K = self.trainer.K
        self.group = group
        self.logit_scale = -math.log(init_temperature)
        self.max_logit_scale = -math.log(min_temperature)
        self.fit_temperature = fit_temperature
        if fit_temperature:
            device = self.trainer.flat.device
            self.temps = torch.full((1,), self.logit_scale, device=device, dtype=torch.float32)
            self.temps_grad, self.temps_m, self.temps_v = (torch.zeros_like(self.temps) for _ in range(3))

    def training_step(self, batch: Mapping[str, Any], _batch_idx: int = 0, optimize: bool = True) -> torch.Tensor:
        K, tr = self.K, self.trainer
        tr.zero_grad()
        v_local = tr.encode_video(batch["video"])
        t_local = tr.encode_text(batch["text"]["input_ids"])
        n = v_local.shape[0]
        if self.fit_temperature:
            self.logit_scale = float(self.temps[0])
        scale = math.exp(self.logit_scale)
        # _step_end (:57-76): all_gather with sync_grads, then the loss on the global batch
        v, off = _all_gather_rows(v_local, self.group)
        t, _ = _all_gather_rows(t_local, self.group)
        scores = K.sgemm(v, t, trans_b=True, alpha=scale)
        loss, dscores = K.loss_fwd_bwd(scores, None, gscale=1.0)
        if self.fit_temperature:  # scores = exp(logit_scale) * V T^T  =>  dL/d logit_scale = sum(dL/dscores * scores)
            self.temps_grad[0] = (dscores * scores).sum()
        tr.backward_text(K.sgemm(dscores[:, off:off + n], v, trans_a=True, alpha=scale))
        tr.backward_video(K.sgemm(dscores[off:off + n], t, alpha=scale))
        if optimize:
            tr.optimizer_step(self.group, extra_grads=[self.temps_grad] if self.fit_temperature else ())
            if self.fit_temperature:
                _adamw_scales_step(tr, self.temps, self.temps_grad, self.temps_m, self.temps_v, self.max_logit_scale)
        return loss


class TeacherStudentTrainingModule:
    """``TeacherStudentLightningModule`` training path without Lightning: :meth:`training_step` is
    ``training_step`` + ``training_step_end`` + ``backward`` + ``optimizer_step`` of the reference loop.

    ``batch``: ``video_student`` / ``video_teacher`` ``(B,T,3,R,R)``, ``text_student`` / ``text_teacher``
    ``{"input_ids": (B, 77)}``, and ``dataset``: a sequence of dataset names, one per sample, grouped
    (``teacher_student.py:101-103``); default: all samples belong to the unlabelled dataset.  ``prompts``: texts that
    replace the unlabelled section's captions in every step (``:104-120``), giving (videos x prompts) score matrices
    there.  ``group``: the process
    group to gather embeddings / reduce gradients over (None = the default group when one is initialised, False = never
    communicate)."""

    def __init__(self, encoder, teacher, init_temperature: float = 0.05, labeled_dataset_name: str = "labeled",
                 labeled_dataset_loss_share: Optional[float] = None,
                 dataset_names: Sequence[str] = ("labeled", "unlabeled"), lr: float = 3e-6,
                 weight_decay: float = 1e-2, group=None, kernels: Any = None, fit_temperature: bool = False,
                 min_temperature: float = 0.001, prompts: Optional[Sequence[str]] = None,
                 gradient_clip_val: Optional[float] = None) -> None:
        self.encoder, self.teacher = encoder, teacher
        self.trainer = ClipTrainer(encoder.model, lr=lr, weight_decay=weight_decay, kernels=kernels,
                                   gradient_clip_val=gradient_clip_val)
        self.K = self.trainer.K
        self.group = group
        self.logit_scale = -math.log(init_temperature)                # video_text_module.py:32
        self.teacher_student_logit_scale = self.logit_scale          # teacher_student.py:70-71
        self.max_logit_scale = -math.log(min_temperature)             # video_text_module.py:34
        # fit_temperature (video_text_module.py:32, teacher_student.py:70-71; config/trainer.yaml:20 defaults to false):
        # the two log-space scales become parameters of the same AdamW group, clamped at max_logit_scale after every
        # optimizer step (video_text_module.py:93-97, teacher_student.py:211-215).  They live in one 2-element device
        # tensor [logit_scale, teacher_student_logit_scale]; their gradients come from the B x B score matrices the
        # step builds anyway, so this path costs one host read of two floats per step.
        self.fit_temperature = fit_temperature
        if fit_temperature:
            device = self.trainer.flat.device
            self.temps = torch.full((2,), self.logit_scale, device=device, dtype=torch.float32)
            self.temps_grad = torch.zeros_like(self.temps)
            self.temps_m = torch.zeros_like(self.temps)
            self.temps_v = torch.zeros_like(self.temps)
        self.labeled_dataset_name = labeled_dataset_name
        names = list(dataset_names)
        if labeled_dataset_loss_share is None:                        # teacher_student.py:60-66
            self.dataset_loss_share = {n: 1 / len(names) for n in names}
        else:
            self.dataset_loss_share = {n: (1 - labeled_dataset_loss_share) / (len(names) - 1) for n in names}
            self.dataset_loss_share[labeled_dataset_name] = labeled_dataset_loss_share
        self.unlabeled_dataset_name = next(n for n in names if n != labeled_dataset_name)
        # prompts (teacher_student.py:47,81-90): a fixed list of texts that stands in for the unlabelled dataset's own
        # captions in every training step; tokenised once, by the student's and by the teacher's tokenizer
        if prompts is None:
            self.tokenized_prompts = self.teacher_tokenized_prompts = None
        else:
            prompts = list(prompts)
            self.tokenized_prompts = dict(encoder.get_tokenizer()(prompts))
            self.teacher_tokenized_prompts = dict(teacher.get_tokenizer()(prompts))

    @staticmethod
    def _splice(text: Mapping[str, torch.Tensor], prompts: Mapping[str, torch.Tensor], lo: int, hi: int):
        """``_replace_in_tokenized_text`` (teacher_student.py:20-40): rows ``[lo, hi)`` of every tensor of ``text`` are
        replaced by the prompt rows; the narrower of the two is right-padded with zeros first."""
        out = {}
        for k, v in text.items():
            new = prompts[k].to(device=v.device, dtype=v.dtype)
            width = max(v.shape[1], new.shape[1])
            v = torch.nn.functional.pad(v, (0, width - v.shape[1]))
            new = torch.nn.functional.pad(new, (0, width - new.shape[1]))
            out[k] = torch.cat((v[:lo], new, v[hi:]))
        return out

    def _sections(self, batch: Mapping[str, Any], n: int) -> List[Tuple[str, int, int]]:
        names = batch.get("dataset")
        if names is None:
            return [(self.unlabeled_dataset_name, 0, n)]
        out, start = [], 0
        for i in range(1, n + 1):
            if i == n or names[i] != names[start]:
                out.append((names[start], start, i))
                start = i
        return out

    def training_step(self, batch: Mapping[str, Any], _batch_idx: int = 0, optimize: bool = True) -> torch.Tensor:
        K, tr = self.K, self.trainer
        tr.zero_grad()
        n_local = batch["video_student"].shape[0]
        sections = self._sections(batch, n_local)
        text_student, text_teacher = batch["text_student"], batch["text_teacher"]
        text_ranges = [(lo, hi) for _, lo, hi in sections]  # rows of the text batch that belong to each section
        if self.tokenized_prompts is not None:  # training_step, teacher_student.py:104-120 and :128-139
            idx = next(i for i, (name, _, _) in enumerate(sections) if name == self.unlabeled_dataset_name)
            _, lo, hi = sections[idx]
            text_student = self._splice(text_student, self.tokenized_prompts, lo, hi)
            text_teacher = self._splice(text_teacher, self.teacher_tokenized_prompts, lo, hi)
            n_prompts = next(iter(self.tokenized_prompts.values())).shape[0]
            shift = n_prompts - (hi - lo)
            text_ranges = [(a, b) if i < idx else (lo, lo + n_prompts) if i == idx else (a + shift, b + shift)
                           for i, (a, b) in enumerate(text_ranges)]
        # _step (teacher_student.py:93-96): student with saved activations, teacher on the evaluation path
        v_local = tr.encode_video(batch["video_student"])
        t_local = tr.encode_text(text_student["input_ids"])
        with torch.no_grad():
            tv_local = self.teacher.encode_video(batch["video_teacher"])
            tt_local = self.teacher.encode_text(text_teacher)
        # _dataset_step_end (:142-173): gather across ranks (sections are per-rank contiguous, so gather per section)
        if self.fit_temperature:
            self.logit_scale, self.teacher_student_logit_scale = self.temps.tolist()
            self.temps_grad.zero_()
        scale, ts_scale = math.exp(self.logit_scale), math.exp(self.teacher_student_logit_scale)
        dv = torch.zeros_like(v_local)
        dt = torch.zeros_like(t_local)
        total = None
        for (name, lo, hi), (tlo, thi) in zip(sections, text_ranges):
            v, off = _all_gather_rows(v_local[lo:hi].contiguous(), self.group)
            t, toff = _all_gather_rows(t_local[tlo:thi].contiguous(), self.group)
            share = self.dataset_loss_share[name]
            scores = K.sgemm(v, t, trans_b=True, alpha=scale)
            if name == self.labeled_dataset_name:
                loss, dscores = K.loss_fwd_bwd(scores, None, gscale=share)
                loss = loss * share
            else:
                tv, _ = _all_gather_rows(tv_local[lo:hi].contiguous(), self.group)
                tt, _ = _all_gather_rows(tt_local[tlo:thi].contiguous(), self.group)
                teacher_scores = K.sgemm(tv, tt, trans_b=True, alpha=ts_scale)
                loss, dscores = K.loss_fwd_bwd(scores, teacher_scores, gscale=share * ts_scale ** 2)
                loss = loss * (share * ts_scale ** 2)
                if self.fit_temperature:
                    self.temps_grad[1] += self._teacher_scale_grad(scores, teacher_scores, loss, share * ts_scale ** 2)
            if self.fit_temperature:  # scores = exp(logit_scale) * V T^T  =>  dL/d logit_scale = sum(dL/dscores * scores)
                self.temps_grad[0] += (dscores * scores).sum()
            total = loss if total is None else total + loss
            # scores = scale * V T^T:  dV = scale * dS T,  dT = scale * dS^T V; keep this rank's rows (videos) / columns
            # (texts: with prompts every rank holds the same list and the gathered matrix carries one copy per rank,
            # exactly as the reference's all_gather of the encoded prompts does, teacher_student.py:144)
            n, nt = hi - lo, thi - tlo
            dv[lo:hi] = K.sgemm(dscores[off:off + n], t, alpha=scale)
            dt[tlo:thi] = K.sgemm(dscores[:, toff:toff + nt], v, trans_a=True, alpha=scale)
        tr.backward_text(dt)
        tr.backward_video(dv)
        if optimize:
            tr.optimizer_step(self.group, extra_grads=[self.temps_grad] if self.fit_temperature else ())
            if self.fit_temperature:
                self._temperature_step()
        return total

    @staticmethod
    def _teacher_scale_grad(scores: torch.Tensor, teacher_scores: torch.Tensor, weighted_loss: torch.Tensor,
                            weight: float) -> torch.Tensor:
        """d/d ts of ``weight(ts) * KL(scores, teacher_scores(ts))`` with ``teacher_scores = exp(ts) * M`` and
        ``weight = share * exp(ts)^2`` (teacher_student.py:156-158): the product rule gives ``2 * weighted_loss`` plus the
        path through the teacher's soft targets, ``weight * sum(dKL/dT * T)`` with, per softmax direction,
        ``dKL/dT = p * (a - sum_line(p * a)) / B``, ``p = softmax(T)``, ``a = log p - log softmax(scores)``."""
        through_targets = scores.new_zeros(())
        for dim in (1, 0):
            B = scores.shape[1 - dim]  # "batchmean": the number of lines of that direction (rows, then columns)
            logp, logq = torch.log_softmax(teacher_scores, dim), torch.log_softmax(scores, dim)
            p, a = logp.exp(), logp - logq
            d = p * (a - (p * a).sum(dim, keepdim=True)) / B
            through_targets = through_targets + (d * teacher_scores).sum()
        return 2 * weighted_loss.reshape(()) + weight * through_targets

    def _temperature_step(self) -> None:
        _adamw_scales_step(self.trainer, self.temps, self.temps_grad, self.temps_m, self.temps_v, self.max_logit_scale)
