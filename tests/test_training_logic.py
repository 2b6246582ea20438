"""Orchestration of the explicit training step (fitclip_b200/training.py) on CPU: with torch
stand-ins for every kernel (tests/torch_kernels.py) the trainer's loss, every parameter gradient and the AdamW update
must equal torch.autograd + torch.optim.AdamW on the oracle's restatement of the reference step
(oracle/train_ref.py <- aligner/teacher_student.py:93-183).  The kernels themselves are checked on the GPU
(tests/test_gpu_train_kernels.py); the two together with tests/test_gpu_training.py cover the CUDA path."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fitclip_b200 import B200ClipVideoTextEncoder
from fitclip_b200.training import ClipTrainer, TeacherStudentTrainingModule
from torch_kernels import TorchKernels

TINY = dict(embed_dim=32, image_resolution=32, vision_layers=2, vision_width=128, vision_patch_size=8,
            context_length=12, vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=2)


def make_models():
    student = oracle.clip_vit_b_16(seed=0, **TINY)
    teacher = oracle.clip_vit_b_16(seed=1, **TINY)
    # non-trivial LayerNorm affines and biases, so that every gradient path carries signal
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in student.named_parameters():
            if n.endswith("bias") or "ln_" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return student.train(), teacher.eval()


def make_batch(n, seed=0, names=None):
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(n, 2, 3, 32, 32, generator=g)
    ids = oracle.tokenize_synthetic(n, (4, 12), seed=seed + 1, context_length=12, vocab_size=64)
    batch = {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids},
             "text_teacher": {"input_ids": ids}}
    if names is not None:
        batch["dataset"] = names
    return batch


def sections_of(names, n):
    if names is None:
        return [("unlabeled", 0, n)]
    out, start = [], 0
    for i in range(1, n + 1):
        if i == n or names[i] != names[start]:
            out.append((names[start], start, i))
            start = i
    return out


@pytest.mark.parametrize("names", [None, ["labeled"] * 3 + ["unlabeled"] * 5])
def test_gradients_and_adamw_match_autograd(names):
    student, teacher = make_models()
    ref_student = oracle.RefClipVideoTextEncoder(copy.deepcopy(student))
    ref_teacher = oracle.RefClipVideoTextEncoder(teacher)
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    module = TeacherStudentTrainingModule(enc, ref_teacher, lr=1e-3, kernels=TorchKernels())
    opt = torch.optim.AdamW(ref_student.model.parameters(), lr=1e-3)
    for step in range(2):
        batch = make_batch(8, seed=10 * step, names=names)
        shares = {"labeled": 0.5, "unlabeled": 0.5}
        ref_loss, ref_grads = oracle.ref_training_step(ref_student, ref_teacher, batch, sections_of(names, 8), opt,
                                                       shares=shares)
        loss = module.training_step(batch, step, optimize=False)
        assert torch.allclose(loss, ref_loss, rtol=1e-4, atol=1e-5), (float(loss), float(ref_loss))
        tr = module.trainer
        assert set(ref_grads) <= set(tr.g)
        for name, ref in ref_grads.items():
            got = tr.g[name]
            scale = ref.abs().max().item()
            err = (got - ref).abs().max().item()
            assert err <= 2e-4 * scale + 1e-6, f"{name}: {err:.3e} vs {scale:.3e}"
        # Adam divides by sqrt(v): entries whose true gradient is zero (the key bias of every attention: softmax is
        # shift-invariant) are pure rounding noise and get a +-lr update of arbitrary sign in BOTH implementations.  The
        # optimizer is therefore compared on identical gradient inputs (their parity is asserted above).
        for name, ref in ref_grads.items():
            tr.g[name].copy_(ref)
        tr.optimizer_step()
        for name, p in ref_student.model.named_parameters():
            assert torch.allclose(tr.w[name], p.detach(), rtol=1e-5, atol=1e-6), name
        # the module's own parameters ARE the trained ones (views into the flat buffer)
        assert enc.model.visual.proj.data_ptr() == tr.w["visual.proj"].data_ptr()


def test_recomputed_layernorm_gives_the_same_gradients():
    student, teacher = make_models()
    grads = []
    for keep in (True, False):
        enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
        module = TeacherStudentTrainingModule(enc, oracle.RefClipVideoTextEncoder(teacher), kernels=TorchKernels())
        module.trainer.keep_layernorm = keep
        module.training_step(make_batch(6, seed=2), 0, optimize=False)
        grads.append(module.trainer.grad.clone())
    assert torch.equal(grads[0], grads[1])


def test_flat_buffers_and_weight_copies():
    student, _ = make_models()
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    tr = ClipTrainer(enc.model, kernels=TorchKernels())
    sd = student.state_dict()
    for name, p in enc.model.named_parameters():
        assert torch.equal(p, sd[name]) and p.data_ptr() >= tr.flat.data_ptr()
        assert (p.data_ptr() - tr.flat.data_ptr()) % 256 == 0
    assert torch.equal(tr.wb["visual.proj"], tr.w["visual.proj"])  # mirror in the activation dtype
    tr.grad.fill_(1.0)
    tr.zero_grad()
    assert float(tr.grad.abs().sum()) == 0.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        student, teacher = make_models()
        enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
        module = TeacherStudentTrainingModule(enc, oracle.RefClipVideoTextEncoder(teacher), lr=1e-3,
                                              kernels=TorchKernels())
        full = make_batch(8, seed=3)
        per = 8 // world
        sl = slice(rank * per, (rank + 1) * per)
        local = {"video_student": full["video_student"][sl], "video_teacher": full["video_teacher"][sl],
                 "text_student": {"input_ids": full["text_student"]["input_ids"][sl]},
                 "text_teacher": {"input_ids": full["text_teacher"]["input_ids"][sl]}}
        loss = module.training_step(local, 0, optimize=True)
        torch.save({"loss": loss, "flat": module.trainer.flat.clone(), "grad": module.trainer.grad.clone()},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_one(tmp_path):
    """Data-parallel step on 2 gloo ranks (embeddings all-gathered, flat gradient all-reduced) == the same step on the
    whole batch in one process."""
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    student, teacher = make_models()
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    module = TeacherStudentTrainingModule(enc, oracle.RefClipVideoTextEncoder(teacher), lr=1e-3, kernels=TorchKernels())
    loss = module.training_step(make_batch(8, seed=3), 0, optimize=True)
    for rank in range(2):
        got = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        assert torch.allclose(got["loss"], loss, rtol=1e-5, atol=1e-6)
        scale = module.trainer.grad.abs().max().item()
        assert (got["grad"] - module.trainer.grad).abs().max().item() <= 1e-4 * scale
        # parameters: compared where the gradient is above rounding noise (Adam turns noise into +-lr steps)
        sig = module.trainer.grad.abs() > 1e-4 * scale
        assert torch.allclose(got["flat"][sig], module.trainer.flat[sig], rtol=1e-4, atol=2e-5)
    a, b = (torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(2))
    assert torch.equal(a["grad"], b["grad"]) and torch.equal(a["flat"], b["flat"])  # replicas stay bit-identical


def test_labeled_dataset_loss_share():
    """``labeled_dataset_loss_share=0.3`` -> shares {labeled: 0.3, unlabeled: 0.7} (teacher_student.py:62-66)."""
    student, teacher = make_models()
    ref_student = oracle.RefClipVideoTextEncoder(copy.deepcopy(student))
    ref_teacher = oracle.RefClipVideoTextEncoder(teacher)
    names = ["labeled"] * 3 + ["unlabeled"] * 5
    batch = make_batch(8, seed=4, names=names)
    opt = torch.optim.AdamW(ref_student.model.parameters(), lr=1e-3)
    ref_loss, ref_grads = oracle.ref_training_step(ref_student, ref_teacher, batch, sections_of(names, 8), opt,
                                                   shares={"labeled": 0.3, "unlabeled": 0.7})
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    module = TeacherStudentTrainingModule(enc, ref_teacher, labeled_dataset_loss_share=0.3, kernels=TorchKernels())
    assert module.dataset_loss_share == {"labeled": 0.3, "unlabeled": 0.7}
    loss = module.training_step(batch, 0, optimize=False)
    assert torch.allclose(loss, ref_loss, rtol=1e-4, atol=1e-5)
    for name, ref in ref_grads.items():
        assert (module.trainer.g[name] - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-6, name


def test_fit_temperature_matches_autograd_and_clamps():
    """``fit_temperature=True`` (video_text_module.py:32, teacher_student.py:70-71): the two log-space scales are
    parameters of the same AdamW group; gradients vs. torch.autograd on the reference's loss expression
    (teacher_student.py:142-173, 176-183), one AdamW step vs. torch.optim.AdamW, then the clamp of ``optimizer_step``
    (video_text_module.py:93-97, teacher_student.py:211-215)."""
    import math
    student, teacher = make_models()
    ref_student = oracle.RefClipVideoTextEncoder(copy.deepcopy(student))
    ref_teacher = oracle.RefClipVideoTextEncoder(teacher)
    names = ["labeled"] * 3 + ["unlabeled"] * 5
    batch = make_batch(8, seed=6, names=names)
    # ---- reference arithmetic with autograd: logit scales as nn.Parameters
    ls = torch.nn.Parameter(torch.tensor([-math.log(0.05)]))
    ts = torch.nn.Parameter(ls.detach().clone())
    params = list(ref_student.model.parameters()) + [ls, ts]
    opt = torch.optim.AdamW(params, lr=1e-3)
    v, t = ref_student(batch["video_student"], batch["text_student"])
    with torch.no_grad():
        tv, tt = ref_teacher(batch["video_teacher"], batch["text_teacher"])
    total = 0.0
    for name, lo, hi in sections_of(names, 8):
        scores = ls.exp() * v[lo:hi] @ t[lo:hi].T
        if name == "labeled":
            loss = oracle.ref_nce_loss(scores)
        else:
            teacher_scores = ts.exp() * tv[lo:hi] @ tt[lo:hi].T
            loss = oracle.ref_teacher_student_nce_loss(scores, teacher_scores, reduction="batchmean") * ts.exp() ** 2
        total = total + loss * 0.5
    total.backward()
    opt.step()
    # ---- the trainer
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    module = TeacherStudentTrainingModule(enc, ref_teacher, lr=1e-3, kernels=TorchKernels(), fit_temperature=True)
    loss = module.training_step(batch, 0)
    assert torch.allclose(loss, total.detach(), rtol=1e-4, atol=1e-5)
    assert abs(float(module.temps_grad[0]) - float(ls.grad)) <= 2e-4 * abs(float(ls.grad)) + 1e-6
    assert abs(float(module.temps_grad[1]) - float(ts.grad)) <= 2e-4 * abs(float(ts.grad)) + 1e-6
    assert abs(float(module.temps[0]) - float(ls)) <= 1e-5 and abs(float(module.temps[1]) - float(ts)) <= 1e-5
    # the next step uses the UPDATED scales
    module.training_step(batch, 1, optimize=False)
    assert abs(module.logit_scale - float(ls)) <= 1e-5 and abs(module.teacher_student_logit_scale - float(ts)) <= 1e-5
    # ---- clamp: a scale pushed past max_logit_scale = -log(min_temperature) comes back to it
    module.temps.fill_(module.max_logit_scale + 1.0)
    module.training_step(batch, 2)
    assert float(module.temps.max()) <= module.max_logit_scale + 1e-6


def test_fit_temperature_gradients_with_prompts_match_autograd():
    """Prompts make the unlabelled score matrices rectangular (5 videos x 3 prompts): ``batchmean`` then divides the row
    direction by 5 and the column direction by 3, which the scale gradients (sum dL/dS * S and the product rule through the
    teacher's soft targets) must follow -- against autograd on the reference's expression (teacher_student.py:104-173)."""
    import math
    student, teacher = make_models()
    ref_student = oracle.RefClipVideoTextEncoder(copy.deepcopy(student))
    ref_teacher = oracle.RefClipVideoTextEncoder(teacher)
    names = ["labeled"] * 3 + ["unlabeled"] * 5
    batch = make_batch(8, seed=9, names=names)
    prompt_ids = make_batch(3, seed=10)["text_student"]["input_ids"]
    ls = torch.nn.Parameter(torch.tensor([-math.log(0.05)]))
    ts = torch.nn.Parameter(ls.detach().clone())
    text = {"input_ids": torch.cat((batch["text_student"]["input_ids"][:3], prompt_ids))}
    v, t = ref_student(batch["video_student"], text)
    with torch.no_grad():
        tv, tt = ref_teacher(batch["video_teacher"], text)
    labeled = oracle.ref_nce_loss(ls.exp() * v[:3] @ t[:3].T)
    scores, teacher_scores = ls.exp() * v[3:] @ t[3:].T, ts.exp() * tv[3:] @ tt[3:].T
    assert scores.shape == (5, 3)
    total = 0.5 * labeled + 0.5 * oracle.ref_teacher_student_nce_loss(scores, teacher_scores, reduction="batchmean") * ts.exp() ** 2
    total.backward()
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    for e in (enc, ref_teacher):
        e.get_tokenizer = lambda: (lambda texts: {"input_ids": prompt_ids.clone()})
    module = TeacherStudentTrainingModule(enc, ref_teacher, lr=1e-3, kernels=TorchKernels(), fit_temperature=True,
                                          prompts=["p0", "p1", "p2"])
    loss = module.training_step(batch, 0, optimize=False)
    assert torch.allclose(loss, total.detach(), rtol=1e-4, atol=1e-5)
    assert abs(float(module.temps_grad[0]) - float(ls.grad)) <= 2e-4 * abs(float(ls.grad)) + 1e-6
    assert abs(float(module.temps_grad[1]) - float(ts.grad)) <= 2e-4 * abs(float(ts.grad)) + 1e-6
    for name, p in ref_student.model.named_parameters():
        if p.grad is not None and name in module.trainer.g:
            err, scale = (module.trainer.g[name] - p.grad).abs().max().item(), p.grad.abs().max().item()
            assert err <= 3e-4 * scale + 1e-6, name


def test_single_encoder_nce_training_matches_autograd():
    """``VideoTextLightningModule`` training (video_text_module.py:25-97): NCE on ``exp(logit_scale) * V @ T.T`` with the
    logit scale as a trained parameter, AdamW, clamp -- against torch.autograd + torch.optim.AdamW on the oracle."""
    import math
    from fitclip_b200.training import VideoTextTrainingModule
    student, _ = make_models()
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(student))
    ls = torch.nn.Parameter(torch.tensor([-math.log(0.05)]))
    opt = torch.optim.AdamW(list(ref.model.parameters()) + [ls], lr=1e-3)
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    module = VideoTextTrainingModule(enc, lr=1e-3, kernels=TorchKernels())
    for step in range(2):
        full = make_batch(6, seed=20 + step)
        batch = {"video": full["video_student"], "text": full["text_student"], "video_id": ["v"] * 6}
        opt.zero_grad(set_to_none=True)
        v, t = ref(batch["video"], batch["text"])
        expect = oracle.ref_nce_loss(ls.exp() * v @ t.T)
        expect.backward()
        loss = module.training_step(batch, step, optimize=False)
        assert torch.allclose(loss, expect.detach(), rtol=1e-4, atol=1e-5)
        assert abs(float(module.temps_grad[0]) - float(ls.grad)) <= 2e-4 * abs(float(ls.grad)) + 1e-6
        for name, p in ref.model.named_parameters():
            if p.grad is not None:
                err, scale = (module.trainer.g[name] - p.grad).abs().max().item(), p.grad.abs().max().item()
                assert err <= 2e-4 * scale + 1e-6, name
        # optimizer on identical gradients (see test_gradients_and_adamw_match_autograd), then the scale's own AdamW step
        for name, p in ref.model.named_parameters():
            if p.grad is not None:
                module.trainer.g[name].copy_(p.grad)
        module.temps_grad[0] = float(ls.grad)
        opt.step()
        module.trainer.optimizer_step()
        from fitclip_b200.training import _adamw_scales_step
        _adamw_scales_step(module.trainer, module.temps, module.temps_grad, module.temps_m, module.temps_v,
                           module.max_logit_scale)
        assert abs(float(module.temps[0]) - float(ls)) <= 1e-5
        for name, p in ref.model.named_parameters():
            assert torch.allclose(module.trainer.w[name], p.detach(), rtol=1e-5, atol=1e-6), name
    # clamp (video_text_module.py:93-97)
    module.temps.fill_(module.max_logit_scale + 2.0)
    module.training_step(batch, 2)
    assert float(module.temps[0]) <= module.max_logit_scale + 1e-6


def test_gradient_clipping_matches_clip_grad_norm():
    """``Trainer(gradient_clip_val=...)`` (config/trainer.yaml:39): global L2 norm over every gradient of the optimizer, the
    trained logit scale included -- against torch.nn.utils.clip_grad_norm_ on the oracle."""
    import math
    from fitclip_b200.training import VideoTextTrainingModule
    student, _ = make_models()
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(student))
    ls = torch.nn.Parameter(torch.tensor([-math.log(0.05)]))
    full = make_batch(6, seed=33)
    batch = {"video": full["video_student"], "text": full["text_student"]}
    v, t = ref(batch["video"], batch["text"])
    oracle.ref_nce_loss(ls.exp() * v @ t.T).backward()
    params = [p for p in ref.model.parameters() if p.grad is not None] + [ls]
    before = torch.sqrt(sum(p.grad.double().pow(2).sum() for p in params))
    clip = 0.25 * float(before)
    torch.nn.utils.clip_grad_norm_(params, clip)
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    module = VideoTextTrainingModule(enc, lr=1e-3, kernels=TorchKernels(), gradient_clip_val=clip)
    module.training_step(batch, 0)  # optimizer_step clips self.grad / temps_grad in place before AdamW
    for name, p in ref.model.named_parameters():
        if p.grad is not None:
            err, scale = (module.trainer.g[name] - p.grad).abs().max().item(), p.grad.abs().max().item()
            assert err <= 3e-4 * scale + 1e-7, name
    assert abs(float(module.temps_grad[0]) - float(ls.grad)) <= 3e-4 * abs(float(ls.grad)) + 1e-7
    after = torch.sqrt(module.trainer.grad.double().pow(2).sum() + module.temps_grad.double().pow(2).sum())
    assert abs(float(after) - clip) <= 1e-3 * clip
    # a threshold above the norm leaves the gradients alone
    enc2 = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2)
    loose = VideoTextTrainingModule(enc2, lr=1e-3, kernels=TorchKernels(), gradient_clip_val=10 * float(before))
    loose.training_step(batch, 0)
    untouched = torch.sqrt(loose.trainer.grad.double().pow(2).sum() + loose.temps_grad.double().pow(2).sum())
    assert abs(float(untouched) - float(before)) <= 1e-3 * float(before)
