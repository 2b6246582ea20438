"""Times the teacher-student TRAINING step at BASELINE configs[4]'s shape on one GPU:
`--videos` videos x 4 frames x 224^2 + as many 77-token captions through the student (forward with saved activations,
backward, AdamW) and the frozen teacher (evaluation path), ViT-B/16, random init.

    python tools/train_step.py --videos 512 --steps 3 --warmup 2 > profiles/r1_train_step_1gpu.json

Prints one JSON line: ms per step, videos/s, the phase split from CUDA events, per-kernel-class times from the
library's per-launch profiler (a second, separately run step), peak memory."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FWD_FLOP_PER_FRAME = 35_126_906_880
FWD_FLOP_PER_CAPTION = 5_959_540_736


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=512)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=12)
    args = ap.parse_args()
    from fitclip_b200 import B200ClipVideoTextEncoder, _lib
    from fitclip_b200._init import init_clip_state_dict
    from fitclip_b200.training import TeacherStudentTrainingModule
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:  # torchrun: data parallel, `--videos` per GPU (weak scaling)
        torch.distributed.init_process_group("nccl", device_id=dev)
    kw = dict(vision_layers=args.layers, transformer_layers=args.layers)
    enc = B200ClipVideoTextEncoder(init_clip_state_dict(seed=0, **kw), num_frames=args.frames).to(dev)
    teach = B200ClipVideoTextEncoder(init_clip_state_dict(seed=1, **kw), num_frames=args.frames).to(dev)
    module = TeacherStudentTrainingModule(enc, teach)
    n = args.videos
    g = torch.Generator(device=dev).manual_seed(1234 + local)
    video = torch.randn(n, args.frames, 3, 224, 224, device=dev, generator=g)
    ids = torch.randint(1, 49405, (n, 77), device=dev, generator=g, dtype=torch.int32)
    ids[:, 0], ids[:, -1] = 49406, 49407
    batch = {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids},
             "text_teacher": {"input_ids": ids}}
    losses = []
    for i in range(args.warmup):
        losses.append(float(module.training_step(batch, i)))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = module.training_step(batch, i)
    e1.record()
    torch.cuda.synchronize()
    losses.append(float(loss))
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:  # the slowest rank sets the step
        t_ms = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
        ms = float(t_ms)
        n_total = n * world
        if torch.distributed.get_rank() == 0:
            print(json.dumps({"metric": "teacher-student training step, data parallel", "n_gpus": world,
                              "videos_per_step": n_total, "ms_per_step": ms, "videos_per_sec": n_total / ms * 1e3,
                              "scaling": "weak", "losses": losses}))
        torch.distributed.destroy_process_group()
        return
    launches = (_lib.launch_count() - launches0) // args.steps

    # phase split of one more step (CUDA events on the same stream)
    tr = module.trainer
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    tr.zero_grad()
    marks[0].record()
    v = tr.encode_video(video)
    marks[1].record()
    t = tr.encode_text(ids)
    marks[2].record()
    with torch.no_grad():
        teach.encode_video(video)
        teach.encode_text({"input_ids": ids})
    marks[3].record()
    tr.backward_text(torch.randn_like(t) * 1e-3)
    marks[4].record()
    tr.backward_video(torch.randn_like(v) * 1e-3)
    marks[5].record()
    tr.optimizer_step()
    marks[6].record()
    torch.cuda.synchronize()
    names = ["student_video_fwd", "student_text_fwd", "teacher_fwd", "text_bwd", "video_bwd", "adamw_and_copies"]
    phases = {k: marks[i].elapsed_time(marks[i + 1]) for i, k in enumerate(names)}

    _lib.profile_start()
    module.training_step(batch, 0)
    recs = _lib.profile_stop()
    kinds = {0: "gemm_tcgen05", 1: "attention", 2: "layernorm", 3: "other"}
    train_tags = {10: "transpose", 11: "colsum", 12: "layernorm_bwd", 13: "quickgelu", 14: "quickgelu_bwd", 15: "adamw"}
    by_kind = {}
    for r in recs:
        key = kinds.get(r["kind"], "other")
        if r["kind"] == 0 and r["tag"] == 9:
            key = "gemm_wgrad_splitk"
        if r["kind"] == 1 and r["tag"] >= 2:
            key = "attention_bwd"
        if r["kind"] == 3 and r["tag"] in train_tags:
            key = train_tags[r["tag"]]
        d = by_kind.setdefault(key, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += r["ms"]
        d["flops"] += r["flops"]
        d["bytes"] += r["bytes"]
        d["launches"] += r["launches"]
    for d in by_kind.values():
        d["tflops"] = d["flops"] / d["ms"] / 1e9 if d["ms"] else 0.0
        d["gbs"] = d["bytes"] / d["ms"] / 1e6 if d["ms"] else 0.0
    fwd = n * args.frames * FWD_FLOP_PER_FRAME + n * FWD_FLOP_PER_CAPTION
    scale = args.layers / 12
    algorithmic = 4 * fwd * scale  # student forward + backward (2x) + teacher forward
    print(json.dumps({
        "metric": "teacher-student training step (ViT-B/16 student fwd+bwd+AdamW, frozen teacher fwd)",
        "videos_per_step": n, "frames_per_video": args.frames, "captions_per_step": n, "layers": args.layers,
        "ms_per_step": ms, "videos_per_sec": n / ms * 1e3, "steps": args.steps, "warmup": args.warmup,
        "algorithmic_tflop_per_step": algorithmic / 1e12, "achieved_tflops": algorithmic / ms / 1e9,
        "gpu_launches_per_step": launches, "phases_ms": phases, "profiled_ms_by_kernel_class": by_kind,
        "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9, "losses": losses,
        "dtype": "bf16 activations / fp32 master weights, gradients and Adam moments"}))


if __name__ == "__main__":
    main()
