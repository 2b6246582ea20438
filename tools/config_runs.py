"""BASELINE.json configs[2] and configs[4] on one GPU (synthetic data, random-init ViT-B/16), one JSON line each:

  configs[2]  WiSE-ensembled student (alpha = 0.5 lerp of two state dicts) zero-shot classification on the UCF101 shape:
              3783 videos x 8 frames against 101 classes x 48 prompt templates (4848 prompts)
  configs[4]  teacher-student scoring pass: frozen teacher + student forward on 512 videos x 4 frames + 512 captions,
              two scaled 512 x 512 score matrices, NCE (labelled) and KL (unlabelled) losses per step"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import B200ClipVideoTextEncoder, wise  # noqa: E402
from fitclip_b200.classification import VideoTextClassificationModule  # noqa: E402
from fitclip_b200.teacher_student import TeacherStudentScoringModule  # noqa: E402

dev = torch.device("cuda:0")
FLOP_FRAME, FLOP_CAPTION = 35_126_906_880, 5_959_540_736


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


with torch.inference_mode():
    a = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0).state_dict(), num_frames=8).to(dev)
    b = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1).state_dict(), num_frames=8).to(dev)

    # ---- configs[2]
    wise_ms, student = timed(lambda: wise(a, b, weight_for_2=0.5))
    n_videos, n_labels, n_templates, T = 3783, 101, 48, 8
    prompts = oracle.tokenize_synthetic(n_labels * n_templates, (6, 20), seed=3)
    module = VideoTextClassificationModule(student, labels=[str(i) for i in range(n_labels)],
                                           templates=["{} %d" % i for i in range(n_templates)],
                                           tokenized_labels={"input_ids": prompts})
    g = torch.Generator(device=dev).manual_seed(7)
    labels = torch.randint(0, n_labels, (n_videos,), device=dev, generator=g)
    chunk = 250
    buf = torch.randn(chunk, T, 3, 224, 224, device=dev, generator=g)

    def classify():
        module.on_validation_start()  # 4848 prompts -> 101 class embeddings (mean over templates)
        for lo in range(0, n_videos, chunk):
            n = min(chunk, n_videos - lo)
            buf.normal_(generator=g)
            module.validation_step({"video": buf[:n], "target": (None, labels[lo:lo + n])})
        return module.validation_epoch_end()

    cls_ms, metrics = timed(classify)
    flops = n_videos * T * FLOP_FRAME + n_labels * n_templates * FLOP_CAPTION
    print(json.dumps({"config": "configs[2] WiSE(0.5) zero-shot classification, UCF101 shape", "wise_lerp_ms": round(wise_ms, 2),
                      "videos": n_videos, "frames_per_video": T, "prompts": n_labels * n_templates,
                      "eval_ms": round(cls_ms, 1), "videos_per_s": round(n_videos / cls_ms * 1e3, 1),
                      "tflops": round(flops / cls_ms / 1e9, 1), "metrics": {k: float(v) for k, v in metrics.items()}}))

    # ---- configs[4]
    B, T = 512, 4
    a.num_frames = b.num_frames = T
    ts = TeacherStudentScoringModule(a, b, init_temperature=0.015).to(dev)
    video = torch.randn(B, T, 3, 224, 224, device=dev, generator=g)
    ids = oracle.tokenize_synthetic(B, 77, seed=5).to(dev)
    batch = {"video_student": video, "text_student": {"input_ids": ids}, "video_teacher": video,
             "text_teacher": {"input_ids": ids}}

    def step():
        out = ts._step(batch)
        return ts._dataset_step_end(out, dataset_name="labeled"), ts._dataset_step_end(out, dataset_name="unlabeled")

    step_ms, (l_nce, l_kl) = timed(step, reps=3)
    flops = 2 * (B * T * FLOP_FRAME + B * FLOP_CAPTION)
    print(json.dumps({"config": "configs[4] teacher-student scoring pass, 512 videos x 4 frames + 512 captions, 2 models",
                      "ms_per_step": round(step_ms, 1), "videos_per_s": round(B / step_ms * 1e3, 1),
                      "tflops": round(flops / step_ms / 1e9, 1), "loss_labeled": float(l_nce), "loss_unlabeled": float(l_kl)}))
