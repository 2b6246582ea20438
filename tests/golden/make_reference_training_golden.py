"""Freezes what the REFERENCE's own training code computes for one teacher-student step, so that the training oracle
(``oracle/train_ref.py``) and the CUDA training path are pinned against reference code rather than against a restatement.

Runs in the build container only (needs ``/root/reference``).  Third-party imports the image lacks are stubbed exactly
as in ``make_reference_golden.py`` (interfaces only); the arithmetic is the reference's:

* ``aligner/teacher_student.py``   ``TeacherStudentLightningModule.training_step`` (grouping by ``batch["dataset"]``,
  ``_step``, ``split_in_collection``) and ``training_step_end`` (``_dataset_step_end`` per dataset, ``dataset_loss_share``)
* ``aligner/loss.py``              ``NCELoss`` / ``TeacherStudentNCELoss("batchmean")``
* ``aligner/encoder/clip_video_text_encoder.py``  the student / teacher wrappers (around the oracle's CLIP)
* ``util/tensor_utils.py``         ``all_gather`` / ``split_in_collection``
then ``loss.backward()`` (torch.autograd) and one ``torch.optim.AdamW(lr=3e-6)`` step, which is what
``config/trainer.yaml:22-24`` + ``aligner/cli.py:126-134`` hand to Lightning.

    python tests/golden/make_reference_training_golden.py   ->  tests/golden/reference_training.pt
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as base  # noqa: E402  (stubs, REFERENCE, ROOT, oracle)

GEOM = dict(embed_dim=64, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=16,
            context_length=16, vocab_size=512, transformer_width=64, transformer_heads=1, transformer_layers=2)


def main() -> None:
    assert os.path.isdir(base.REFERENCE), f"{base.REFERENCE} is not mounted: this script only runs in the build container"
    torch.set_num_threads(1)
    base.install_stubs()
    sys.path.insert(0, base.REFERENCE)
    from aligner.encoder.clip_video_text_encoder import ClipVideoTextEncoder  # noqa: E402
    from aligner.teacher_student import TeacherStudentLightningModule  # noqa: E402

    oracle = base.oracle
    student = oracle.clip_vit_b_16(seed=0, **GEOM)
    teacher = oracle.clip_vit_b_16(seed=1, **GEOM)
    g = torch.Generator().manual_seed(20221119)
    with torch.no_grad():  # non-trivial LayerNorm affines / biases so that every gradient path carries signal
        for n, p in student.named_parameters():
            if n.endswith("bias") or "ln_" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    out = {"config": GEOM, "student_state_dict": {k: v.clone() for k, v in student.state_dict().items()},
           "teacher_state_dict": {k: v.clone() for k, v in teacher.state_dict().items()},
           "reference_files": ["aligner/teacher_student.py", "aligner/loss.py", "aligner/video_text_module.py",
                               "aligner/text_video_retrieval.py", "aligner/encoder/clip_video_text_encoder.py",
                               "util/tensor_utils.py"]}
    n = 10
    video = torch.randn(n, 2, 3, 32, 32, generator=g)
    ids = oracle.tokenize_synthetic(n, (3, 16), seed=78, context_length=16, vocab_size=512)
    names = ["labeled"] * 4 + ["unlabeled"] * 6
    out.update(video=video, input_ids=ids, dataset=names, init_temperature=0.05)

    enc_s = ClipVideoTextEncoder(student, num_frames=2).train()
    enc_t = ClipVideoTextEncoder(teacher, num_frames=2).eval()
    module = TeacherStudentLightningModule(encoder=enc_s, teacher=enc_t, init_temperature=0.05, fit_temperature=False)
    optimizer = torch.optim.AdamW([p for p in module.parameters() if p.requires_grad], lr=3e-6)
    module.trainer = types.SimpleNamespace(optimizers=[optimizer])
    batch = {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids.clone()},
             "text_teacher": {"input_ids": ids.clone()}, "dataset": list(names)}
    loss = module.training_step_end(module.training_step(batch, 0))
    loss.backward()
    out["loss"] = loss.detach().clone()
    out["logged"] = {name: (v.clone() if isinstance(v, torch.Tensor) else v) for name, v, _ in module.logged
                     if name.startswith("loss/")}
    out["grads"] = {k: p.grad.detach().clone() for k, p in enc_s.model.named_parameters() if p.grad is not None}
    out["trainable"] = sorted(k for k, p in module.named_parameters() if p.requires_grad)
    optimizer.step()
    # the updated parameters of a representative subset (every kind of tensor; the full set would double the fixture)
    keep = ("visual.proj", "visual.class_embedding", "visual.ln_pre.weight", "visual.transformer.resblocks.1.mlp.c_fc.weight",
            "visual.transformer.resblocks.0.attn.in_proj_bias", "transformer.resblocks.1.attn.out_proj.weight",
            "transformer.resblocks.0.ln_2.bias", "positional_embedding", "text_projection", "ln_final.weight")
    after = enc_s.model.state_dict()
    out["student_after_step"] = {k: after[k].detach().clone() for k in keep}
    path = os.path.join(base.ROOT, "tests", "golden", "reference_training.pt")
    torch.save(out, path)

    # ---- the same step with `prompts` (teacher_student.py:47,81-90,104-120,128-139): the unlabelled section's captions are
    # replaced by a fixed prompt list, its score matrices become (6 videos x 5 prompts).  Same weights and inputs as above
    # (reloaded: the optimizer step above moved the student), so the second fixture stores only what differs.
    student.load_state_dict({k: v for k, v in out["student_state_dict"].items() if k != "logit_scale"})  # dropped by the wrapper
    for p in student.parameters():
        p.grad = None  # the first step's gradients are still attached to the same parameter objects
    prompts = ["a video of a person cooking", "a video of a dog", "someone playing guitar", "a cartoon", "news"]
    enc_s = ClipVideoTextEncoder(student, num_frames=2).train()
    module = TeacherStudentLightningModule(encoder=enc_s, teacher=enc_t, init_temperature=0.05, fit_temperature=False,
                                           prompts=prompts)
    module.trainer = types.SimpleNamespace(optimizers=[torch.optim.AdamW(
        [p for p in module.parameters() if p.requires_grad], lr=3e-6)])
    batch = {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids.clone()},
             "text_teacher": {"input_ids": ids.clone()}, "dataset": list(names)}
    loss_p = module.training_step_end(module.training_step(batch, 0))
    loss_p.backward()
    grads = {k: p.grad.detach().clone() for k, p in enc_s.model.named_parameters() if p.grad is not None}
    keep_g = keep + ("visual.conv1.weight", "token_embedding.weight", "transformer.resblocks.0.attn.in_proj_weight",
                     "visual.transformer.resblocks.0.ln_1.weight", "visual.positional_embedding", "ln_final.bias")
    out_p = {"prompts": prompts, "tokenized_prompts": {k: v.detach().clone() for k, v in module.tokenized_prompts.items()},
             "loss": loss_p.detach().clone(), "num_grads": len(grads),
             "logged": {name: (v.clone() if isinstance(v, torch.Tensor) else v) for name, v, _ in module.logged
                        if name.startswith("loss/")},
             "grads": {k: grads[k] for k in keep_g},
             "reference_files": ["aligner/teacher_student.py", "aligner/loss.py", "util/tensor_utils.py"]}
    path_p = os.path.join(base.ROOT, "tests", "golden", "reference_training_prompts.pt")
    torch.save(out_p, path_p)
    print(f"wrote {path_p} ({os.path.getsize(path_p) / 1e6:.2f} MB); loss with prompts {float(loss_p):.6f}")
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB); loss {float(loss):.6f}; {len(out['grads'])} gradient "
          f"tensors; torch {torch.__version__}")


if __name__ == "__main__":
    main()
