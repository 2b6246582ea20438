"""The training step on the CUDA path (fitclip_b200/training.py + csrc/train.cu) against torch.autograd on the oracle's
restatement of the reference step (oracle/train_ref.py <- aligner/teacher_student.py:93-183), same seeded inputs.

Tolerance: the CUDA path keeps activations and activation gradients in bf16 (fp32 accumulation, fp32 parameter
gradients), the oracle is fp32 throughout.  Per parameter tensor: cosine similarity >= 0.98 and norm within 6 %, for
every tensor whose reference gradient is above rounding noise; loss within 2 %."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

GEOM = dict(vision_layers=2, transformer_layers=2, image_resolution=64, context_length=24, vocab_size=512)


def _batch(n, frames, res, ctx, vocab, dev, names=None, seed=0):
    import oracle
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(n, frames, 3, res, res, generator=g).to(dev)
    ids = oracle.tokenize_synthetic(n, (5, ctx), seed=seed + 1, context_length=ctx, vocab_size=vocab).to(dev)
    # dev = "cpu" gives the oracle's copy of the same batch
    batch = {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids},
             "text_teacher": {"input_ids": ids}}
    if names is not None:
        batch["dataset"] = names
    return batch


@pytest.mark.parametrize("names", [None, ["labeled"] * 6 + ["unlabeled"] * 10])
def test_training_step_matches_autograd(dev, names):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    from fitclip_b200.training import TeacherStudentTrainingModule
    student = oracle.clip_vit_b_16(seed=0, **GEOM)
    teacher = oracle.clip_vit_b_16(seed=1, **GEOM)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in student.named_parameters():
            if n.endswith("bias") or "ln_" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    ref_student = oracle.RefClipVideoTextEncoder(copy.deepcopy(student)).train()  # the oracle runs on the CPU
    ref_teacher = oracle.RefClipVideoTextEncoder(copy.deepcopy(teacher)).eval()
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=2).to(dev)
    teach = B200ClipVideoTextEncoder(teacher.state_dict(), num_frames=2).to(dev)
    module = TeacherStudentTrainingModule(enc, teach, lr=1e-4)
    n = 16
    batch = _batch(n, 2, 64, 24, 512, dev, names)
    cpu_batch = _batch(n, 2, 64, 24, 512, "cpu", names)
    sections = [("unlabeled", 0, n)] if names is None else [("labeled", 0, 6), ("unlabeled", 6, n)]
    opt = torch.optim.AdamW(ref_student.model.parameters(), lr=1e-4)
    ref_loss, ref_grads = oracle.ref_training_step(ref_student, ref_teacher, cpu_batch, sections, opt,
                                                   shares={"labeled": 0.5, "unlabeled": 0.5})
    launches0 = __import__("fitclip_b200")._lib.launch_count()
    loss = module.training_step(batch, 0, optimize=False)
    assert __import__("fitclip_b200")._lib.launch_count() - launches0 > 100  # the native kernels ran
    assert abs(float(loss) - float(ref_loss)) <= 0.02 * abs(float(ref_loss)) + 1e-4, (float(loss), float(ref_loss))
    tr = module.trainer
    top = max(float(v.norm()) for v in ref_grads.values())
    checked = 0
    for name, ref in ref_grads.items():
        rn = float(ref.norm())
        if rn < 1e-3 * top:
            continue  # e.g. the key bias of every attention (exactly zero gradient; rounding noise on both sides)
        got = tr.g[name].cpu()
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
        ratio = float(got.norm()) / rn
        assert cos >= 0.995 and abs(ratio - 1) <= 0.08, f"{name}: cos {cos:.4f} norm ratio {ratio:.4f}"
        checked += 1
    assert checked >= 40
    # optimizer on identical gradients (Adam amplifies rounding noise of near-zero gradients into +-lr steps)
    for name, ref in ref_grads.items():
        tr.g[name].copy_(ref)
    tr.optimizer_step()
    for name, p in ref_student.model.named_parameters():
        assert torch.allclose(tr.w[name].cpu(), p.detach(), rtol=1e-5, atol=1e-6), name
    # the evaluation path of the same module now encodes with the UPDATED weights
    with torch.inference_mode():
        v_new = enc.encode_video(batch["video_student"])
        v_ref = ref_student.encode_video(cpu_batch["video_student"])
    cos = torch.nn.functional.cosine_similarity(v_new.cpu(), v_ref, dim=-1)
    assert float(cos.min()) >= 0.999


def test_other_geometry_and_recomputed_layernorm(dev):
    """ViT-B/32-style patches (5 image tokens: one attention tile), the real 77-token context, LayerNorm outputs
    recomputed in the backward pass instead of kept: gradients against autograd on the oracle."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    from fitclip_b200.training import TeacherStudentTrainingModule
    geom = dict(vision_layers=1, transformer_layers=1, image_resolution=64, vision_patch_size=32, context_length=77,
                vocab_size=512)
    student, teacher = oracle.clip_vit_b_16(seed=0, **geom), oracle.clip_vit_b_16(seed=1, **geom)
    ref_student = oracle.RefClipVideoTextEncoder(copy.deepcopy(student)).train()
    ref_teacher = oracle.RefClipVideoTextEncoder(copy.deepcopy(teacher)).eval()
    enc = B200ClipVideoTextEncoder(student.state_dict(), num_frames=3).to(dev)
    teach = B200ClipVideoTextEncoder(teacher.state_dict(), num_frames=3).to(dev)
    module = TeacherStudentTrainingModule(enc, teach)
    module.trainer.keep_layernorm = False
    n = 9
    batch, cpu_batch = _batch(n, 3, 64, 77, 512, dev), _batch(n, 3, 64, 77, 512, "cpu")
    opt = torch.optim.AdamW(ref_student.model.parameters(), lr=3e-6)
    ref_loss, ref_grads = oracle.ref_training_step(ref_student, ref_teacher, cpu_batch, [("unlabeled", 0, n)], opt)
    loss = module.training_step(batch, 0, optimize=False)
    assert abs(float(loss) - float(ref_loss)) <= 0.02 * abs(float(ref_loss)) + 1e-4
    top = max(float(v.norm()) for v in ref_grads.values())
    checked = 0
    for name, ref in ref_grads.items():
        if float(ref.norm()) < 1e-3 * top:
            continue
        got = module.trainer.g[name].cpu()
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
        assert cos >= 0.995 and abs(float(got.norm()) / float(ref.norm()) - 1) <= 0.08, f"{name}: cos {cos:.4f}"
        checked += 1
    assert checked >= 20


def test_native_errors_are_raised_not_swallowed(dev):
    """Bad arguments come back as FitclipError with the library's message (no silent fallback)."""
    from fitclip_b200 import _lib, train_ops as T
    x = torch.zeros(8, 12, device=dev, dtype=torch.bfloat16)  # 12 columns: not a multiple of 8
    with pytest.raises(_lib.FitclipError, match="multiples of 8"):
        T.colsum(x, torch.zeros(12, device=dev))
    with pytest.raises(_lib.FitclipError, match="unsupported"):
        T.attention_bwd(torch.zeros(500, 192, device=dev, dtype=torch.bfloat16),
                        torch.zeros(500, 64, device=dev, dtype=torch.bfloat16),
                        torch.zeros(500, 64, device=dev, dtype=torch.bfloat16), 1, 500, 1, False)
    with pytest.raises(_lib.FitclipError):
        T.colsum(torch.zeros(8, 16, dtype=torch.bfloat16), torch.zeros(16))  # CPU tensors: no CPU path


def test_loss_decreases(dev):
    """A few steps at a large learning rate on one fixed batch: the distillation loss goes down."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    from fitclip_b200.training import TeacherStudentTrainingModule
    enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **GEOM).state_dict(), num_frames=2).to(dev)
    teach = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1, **GEOM).state_dict(), num_frames=2).to(dev)
    module = TeacherStudentTrainingModule(enc, teach, lr=2e-4)
    batch = _batch(16, 2, 64, 24, 512, dev)
    losses = [float(module.training_step(batch, i)) for i in range(6)]
    assert losses[-1] < losses[0], losses


def test_out_of_range_token_ids_are_reported(dev):
    """a bad token id on the TRAINING path (text_embed clamps it, token_scatter drops its gradient) must
    surface through ClipTrainer.check_inputs -- the reference's embedding lookup would have raised."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, _lib
    from fitclip_b200.training import TeacherStudentTrainingModule
    enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **GEOM).state_dict(), num_frames=2).to(dev)
    teach = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1, **GEOM).state_dict(), num_frames=2).to(dev)
    module = TeacherStudentTrainingModule(enc, teach)
    batch = _batch(8, 2, 64, 24, 512, dev)
    module.training_step(batch, 0)
    module.trainer.check_inputs()  # clean batch: no error
    bad = batch["text_student"]["input_ids"].clone()
    bad[3, 2] = 60000
    batch["text_student"] = {"input_ids": bad}  # the teacher keeps the clean ids
    module.training_step(batch, 1)
    with pytest.raises(_lib.FitclipError, match="token id out of range"):
        module.trainer.check_inputs()
    module.trainer.check_inputs()  # the flag was reset
