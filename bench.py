#!/usr/bin/env python
"""Benchmark of the FitCLIP evaluation hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one complete retrieval evaluation of the synthetic MSR-VTT-1kA-shaped workload (BASELINE.json configs[1]):
per GPU 1000 videos x 4 frames x 3x224x224 fp32 + 1000 captions x 77 tokens -> CLIP ViT-B/16 encode (bf16 tensor
cores, random-init weights) -> per-frame L2 norm + frame mean-pool -> text x video similarity -> ranks -> R@1/5/10/MdR.
With N > 1 every rank owns 1000 videos/captions of an N*1000 gallery (weak scaling): video-sharded encode, NCCL
all-gather of text embeddings, column-sharded fused similarity+rank count, all-reduce of target scores and counts.

`value`   videos/s, inputs resident in HBM when the timed region starts (inputs are 2.4 GB/GPU > the 126 MB L2).
`e2e`     videos/s through the public plugin API from pinned HOST buffers (H2D of every frame/token batch and the
          D2H read of the metrics are inside the timed region).
`roofline` the tcgen05 GEMM kernel, timed per launch with CUDA events on its stream during the timed steps.
`cpu_baseline` the oracle (restated reference path, fp32 torch CPU) on a bounded sample, on this box's host cores.

`oracle/` is imported in exactly two places: the CPU legs (`cpu_baseline`, `--impl reference`) and, as a checker outside every
timed region, the sampled-row comparison of the 100k-video leg.  Weights and inputs of the CUDA arm come from the package's
own initialiser and from this file (`synthetic_weights`, `synthetic_tokens`, `synthetic_inputs`); the CPU legs load the same.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME = 35_126_906_880      # SURVEY.md 8d / BASELINE.md section 2
FLOP_PER_CAPTION = 5_959_540_736
VIDEOS_PER_GPU, FRAMES, CAPTIONS_PER_GPU, CTX = 1000, 4, 1000, 77
WORKLOAD = "msrvtt_1ka_shape: 1000 videos x 4 frames x 3x224x224 fp32 + 1000 captions x 77 tok per GPU, ViT-B/16"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"  # B200_PROFILING.md fallback figures


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.rows = []
        self.gpu_index = gpu_index
        self.proc = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, flag in zip(names, r[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_weights(seed: int = 0):
    """The benchmark's weights (no network, no checkpoints): the package's own seeded CLIP initialiser
    (config/encoder/clip_from_scratch_vit_b_16.yaml) with trained-like values wherever an init leaves identities --
    LayerNorm gamma ~ U(0.2, 3), beta ~ N(0, 0.5), attention biases ~ N(0, 0.1) -- so that the folded-LayerNorm arithmetic
    runs on real numbers.  Both arms (ours, and the CPU oracle of `cpu_baseline` / `--impl reference`) load this state dict."""
    import torch

    from fitclip_b200._init import init_clip_state_dict
    sd = init_clip_state_dict(seed=seed)
    g = torch.Generator().manual_seed(1_000_003 + seed)
    for name, p in sd.items():
        parent, _, leaf = name.rpartition(".")
        is_ln = parent.rsplit(".", 1)[-1] in ("ln_1", "ln_2", "ln_pre", "ln_post", "ln_final")
        if is_ln and leaf == "weight":
            p.copy_(0.2 + 2.8 * torch.rand(p.shape, generator=g))
        elif is_ln and leaf == "bias":
            p.copy_(0.5 * torch.randn(p.shape, generator=g))
        elif name.endswith("attn.in_proj_bias") or name.endswith("attn.out_proj.bias"):
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return sd


def synthetic_tokens(count: int, seed: int):
    """SURVEY.md 8d captions: ``[SOT] + random ids + [EOT]``, dense (all CTX positions used), int32."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, 49406, (count, CTX), generator=g, dtype=torch.int32)
    ids[:, 0], ids[:, -1] = 49406, 49407
    return ids


def synthetic_inputs(rank: int, device, pinned: bool):
    """frames ~ N(0,1) (post-Normalize statistics), captions = [SOT] + random ids + [EOT], dense 77."""
    import torch

    g = torch.Generator().manual_seed(1234 + rank)
    ids = synthetic_tokens(CAPTIONS_PER_GPU, 4321 + rank)
    if pinned:
        frames = torch.empty(VIDEOS_PER_GPU, FRAMES, 3, 224, 224, dtype=torch.float32, pin_memory=True)
        frames.normal_(generator=g)
        return frames, ids.pin_memory()
    gd = torch.Generator(device=device).manual_seed(1234 + rank)
    frames = torch.randn(VIDEOS_PER_GPU, FRAMES, 3, 224, 224, device=device, generator=gd)
    return frames, ids.to(device)


WEBVID_FRAMES = 8


def rank_equality_check(encoder, frames, ids, n_total, world, rank, group, device) -> dict:
    """N > 1: the sharded `retrieval_ranks` (column-sharded similarity, all-reduced target scores and counts) against the
    single-GPU evaluation of the SAME embeddings, on every rank: for the even N*1000 split the bench times and for an
    uneven N*1000 + 1 split (`shard_bounds`: the last shard is short)."""
    import torch
    import torch.distributed as dist
    from fitclip_b200 import retrieval_ranks, shard_bounds
    from fitclip_b200.retrieval import NO_GROUP, all_gather_rows
    v = encoder.encode_video(frames)
    t = encoder.encode_text({"input_ids": ids})
    sharded = retrieval_ranks(t, v, group=group, totals=(n_total, n_total))
    v_all, _ = all_gather_rows(v, group, total=n_total)
    t_all, _ = all_gather_rows(t, group, total=n_total)
    single = retrieval_ranks(t_all.contiguous(), v_all.contiguous(), group=NO_GROUP)
    equal = bool(torch.equal(sharded, single))
    # uneven split: one more (video, caption) pair, identical on every rank
    g = torch.Generator().manual_seed(99)
    extra = torch.nn.functional.normalize(torch.randn(2, v.shape[1], generator=g), dim=-1).to(device)
    v2, t2 = torch.cat([v_all, extra[:1]]), torch.cat([t_all, extra[1:]])
    lo, hi = shard_bounds(n_total + 1, world, rank)
    sharded2 = retrieval_ranks(t2[lo:hi].contiguous(), v2[lo:hi].contiguous(), group=group,
                               totals=(n_total + 1, n_total + 1))
    single2 = retrieval_ranks(t2.contiguous(), v2.contiguous(), group=NO_GROUP)
    equal2 = bool(torch.equal(sharded2, single2))
    flags = torch.tensor([int(equal), int(equal2)], device=device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    return {"equal": bool(flags[0].item()), "equal_uneven": bool(flags[1].item()), "gallery": n_total,
            "gallery_uneven": n_total + 1, "checked_on": "every rank (all-reduce MIN of the per-rank verdicts)"}


def webvid_leg(encoder, device, rank, world, group, n_total, barrier) -> dict:
    """One full evaluation of BASELINE configs[3]: `n_total` videos x 8 frames + `n_total` captions SPLIT over the
    ranks (`shard_bounds`: strong scaling), real encoder outputs fed to the sharded similarity + rank count, R@k/MdR.
    Frames cannot be resident (100k x 8 x 3x224x224 fp32 = 482 GB): each chunk of 500 videos is produced on the device
    inside the timed region as an affine map `a_c * pool + b_c` of a resident 500-video N(0,1) pool (one elementwise
    pass, ~1 % of the chunk's encode time), so every video is distinct; token ids are resident.  Afterwards (untimed)
    sampled query rows are ranked by the CPU oracle on the kernel's own scores and compared."""
    import torch

    import oracle
    from fitclip_b200 import metrics_from_ranks, ops, retrieval_ranks, shard_bounds
    from fitclip_b200.retrieval import all_gather_rows
    lo, hi = shard_bounds(n_total, world, rank)
    n_local = hi - lo
    chunk = 500
    gd = torch.Generator(device=device).manual_seed(777)
    pool = torch.randn(chunk, WEBVID_FRAMES, 3, 224, 224, device=device, generator=gd)
    buf = torch.empty_like(pool)
    ids = synthetic_tokens(1000, 555).to(device)  # 1000 captions, varied per block below
    v_emb = torch.empty(n_local, 512, device=device)
    t_emb = torch.empty(n_local, 512, device=device)
    e0, e1, e_enc = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    with torch.inference_mode():
        barrier()
        e0.record()
        for c in range(lo, hi, chunk):
            nb = min(chunk, hi - c)
            k = c // chunk  # GLOBAL chunk index: video i is the same tensor whatever the number of ranks
            shift = torch.full((), ((k * 53) % 97) / 97.0 - 0.5, device=device)
            torch.add(shift, pool[:nb], alpha=0.75 + 0.5 * ((k * 37) % 101) / 101.0, out=buf[:nb])  # one pass
            v_emb[c - lo:c - lo + nb] = encoder.encode_video(buf[:nb])
        for c in range(lo, hi, 1000):
            nb = min(1000, hi - c)
            # caption i = synthetic caption i % 1000 with its body tokens shifted by 7 * (i // 1000) (SOT / EOT stay):
            # distinct captions, and caption i is the same whatever the number of ranks
            idx = torch.arange(c, c + nb, device=device)
            block = ids[idx % 1000]
            block[:, 1:CTX - 1] = ((block[:, 1:CTX - 1] + (idx // 1000 * 7).to(torch.int32).unsqueeze(1)) % 49000 + 1)
            t_emb[c - lo:c - lo + nb] = encoder.encode_text({"input_ids": block})
        e_enc.record()
        ranks = retrieval_ranks(t_emb, v_emb, group=group, totals=(n_total, n_total))
        m = metrics_from_ranks(ranks, n_total)
        m_host = {k: float(x) for k, x in m.items()}  # D2H read of the result inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e_enc)], device=device, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        seconds, enc_seconds = ms[0].item() / 1e3, ms[1].item() / 1e3
        # ---- untimed check: sampled rows, oracle ranking of the kernel's own S ----
        v_all, _ = all_gather_rows(v_emb, group, total=n_total)
        t_all, _ = all_gather_rows(t_emb, group, total=n_total)
        gs = torch.Generator().manual_seed(4)
        rows = torch.randperm(n_total, generator=gs)[:64].sort().values
        scores = ops.Similarity(t_all[rows.to(device)].contiguous(), v_all.contiguous(), 3).scores().cpu()
        expect = oracle.ref_stable_rank(scores, rows)
        ok = bool(torch.equal(ranks[rows.to(device)].cpu(), expect))
    return {"workload": f"webvid_shape: {n_total} videos x {WEBVID_FRAMES} frames x 3x224x224 fp32 + {n_total} captions x 77 "
                        f"tok, split over {world} GPU(s) (strong scaling)", "videos": n_total, "n_gpus": world,
            "seconds": seconds, "encode_seconds": enc_seconds, "sim_rank_seconds": seconds - enc_seconds,
            "videos_per_s": n_total / seconds, "queries_per_s": n_total / seconds, "metrics": m_host,
            "sampled_rows_match_oracle": ok, "sampled_rows": 64,
            "inputs": "500-video resident pool, per-chunk affine map inside the timed region; resident token ids"}


def train_leg(teacher, device, videos: int, steps: int = 3) -> dict:
    """BASELINE configs[4] as a TRAINING step: `videos` videos x 4 frames + as many captions per
    step; student ViT-B/16 forward with saved activations + backward + AdamW, frozen teacher forward (the evaluation
    engine the bench already holds).  One warm-up step, `steps` timed ones, CUDA events; N = 1 only."""
    import torch

    from fitclip_b200 import B200ClipVideoTextEncoder, _lib
    from fitclip_b200._init import init_clip_state_dict
    from fitclip_b200.training import TeacherStudentTrainingModule
    student = B200ClipVideoTextEncoder(init_clip_state_dict(seed=3), num_frames=FRAMES).to(device)
    module = TeacherStudentTrainingModule(student, teacher)
    g = torch.Generator(device=device).manual_seed(99)
    video = torch.randn(videos, FRAMES, 3, 224, 224, device=device, generator=g)
    ids = torch.randint(1, 49405, (videos, CTX), device=device, generator=g, dtype=torch.int32)
    ids[:, 0], ids[:, -1] = 49406, 49407
    batch = {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids},
             "text_teacher": {"input_ids": ids}}
    torch.cuda.reset_peak_memory_stats(device)
    losses = [float(module.training_step(batch, 0))]
    torch.cuda.synchronize(device)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = module.training_step(batch, i + 1)
    e1.record()
    torch.cuda.synchronize(device)
    module.trainer.check_inputs()
    losses.append(float(loss))
    ms = e0.elapsed_time(e1) / steps
    fwd = videos * FRAMES * FLOP_PER_FRAME + videos * FLOP_PER_CAPTION
    algorithmic = 4 * fwd  # teacher forward + student forward + student backward (2x)
    return {"workload": f"teacher-student training step: {videos} videos x {FRAMES} frames + {videos} captions, student "
                        f"fwd + bwd + AdamW and frozen-teacher fwd, ViT-B/16", "steps": steps, "ms_per_step": ms,
            "videos_per_s": videos / ms * 1e3, "algorithmic_tflop_per_step": algorithmic / 1e12,
            "achieved_tflops": algorithmic / ms / 1e9, "gpu_launches_per_step": (_lib.launch_count() - launches0) // steps,
            "peak_memory_gb": torch.cuda.max_memory_allocated(device) / 1e9, "losses_first_last": losses}


def geometry_leg(device, videos: int, steps: int = 3) -> list:
    """SURVEY.md 8f row f4 on the driver's box: the other tower geometries the reference ships configs for, random-init,
    `videos` videos x 4 frames + as many captions -> similarity + ranks per step (the cfg2 step at another shape):
    `clip_vit_l_14` (config/encoder/clip_vit_l_14.yaml: 257 image tokens -> attention_tc_long_kernel) and the SLIP layout
    (config/encoder/slip_vit_b_16.yaml: timm image tower -- no ln_pre, erf-GELU epilogue).  FLOPs: dense algorithmic count."""
    import torch

    from fitclip_b200 import B200ClipVideoTextEncoder, B200SlipVideoTextEncoder, metrics_from_ranks, retrieval_ranks
    from fitclip_b200._init import init_clip_state_dict, init_slip_state_dict

    def tower(tokens, width, layers):
        return layers * (2 * tokens * 12 * width * width + 4 * tokens * tokens * width)

    out = []
    for name, build, kw in (
            ("clip_vit_l_14", lambda sd: B200ClipVideoTextEncoder(sd, num_frames=FRAMES), dict(
                init=init_clip_state_dict, embed_dim=768, vision_patch_size=14, vision_width=1024, vision_layers=24,
                transformer_width=768, transformer_heads=12)),
            ("slip_vit_b_16", lambda sd: B200SlipVideoTextEncoder(sd, num_frames=FRAMES), dict(init=init_slip_state_dict))):
        init = kw.pop("init")
        enc = build(init(seed=11, **kw)).to(device)
        cfg = enc.model.config
        L = (cfg["image_resolution"] // cfg["vision_patch_size"]) ** 2 + 1
        g = torch.Generator(device=device).manual_seed(5)
        video = torch.randn(videos, FRAMES, 3, 224, 224, device=device, generator=g)
        ids = torch.randint(1, 49405, (videos, CTX), device=device, generator=g, dtype=torch.int32)
        ids[:, 0], ids[:, -1] = 49406, 49407

        def step():
            v = enc.encode_video(video)
            t = enc.encode_text({"input_ids": ids})
            return metrics_from_ranks(retrieval_ranks(t, v, group=False), videos)

        with torch.inference_mode():
            step()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / steps
        frame = tower(L, cfg["vision_width"], cfg["vision_layers"]) \
            + 2 * (L - 1) * cfg["vision_width"] * 3 * cfg["vision_patch_size"] ** 2 + 2 * cfg["vision_width"] * cfg["embed_dim"]
        cap = tower(CTX, cfg["transformer_width"], cfg["transformer_layers"]) + 2 * cfg["transformer_width"] * cfg["embed_dim"]
        flops = videos * (FRAMES * frame + cap)
        out.append({"geometry": name, "image_tokens": L, "videos": videos, "frames_per_video": FRAMES, "steps": steps,
                    "ms_per_step": ms, "videos_per_s": videos / ms * 1e3, "achieved_tflops": flops / ms / 1e9,
                    "gflop_per_frame": frame / 1e9})
        del enc, video
        torch.cuda.empty_cache()
    return out


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from fitclip_b200 import B200ClipVideoTextEncoder, _lib, metrics_from_ranks, retrieval_ranks

    # Libraries (NCCL's version banner, ...) may write to fd 1; the contract is ONE JSON line on stdout, so everything
    # but the final print goes to stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"

    # random-init ViT-B/16, identical on every rank (config/encoder/clip_from_scratch_vit_b_16.yaml)
    from fitclip_b200 import B200Clip
    encoder = B200ClipVideoTextEncoder(B200Clip(synthetic_weights(0), max_frames_per_pass=args.frames_per_pass),
                                       num_frames=FRAMES).to(device)
    frames, ids = synthetic_inputs(rank, device, pinned=False)
    n_total = VIDEOS_PER_GPU * world

    def step_resident():
        v = encoder.encode_video(frames)
        t = encoder.encode_text({"input_ids": ids})
        ranks = retrieval_ranks(t, v, group=group, totals=(n_total, n_total))  # analytic shard sizes: no host sync
        return metrics_from_ranks(ranks, n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.inference_mode():
        for _ in range(max(args.warmup, 3)):
            metrics = step_resident()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            metrics = step_resident()
        e1.record()
        barrier()
        launches = _lib.launch_count() - launches0
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_per_step = ms.item() / args.steps
        # The same K steps again with a CUDA-event pair around every launch of the library (per-kernel durations for
        # the roofline).  Kept apart from the run above because an event record between two kernels serialises them,
        # which would switch off the programmatic-dependent-launch overlap the headline number is entitled to.
        _lib.profile_start(1 << 16)
        barrier()
        e0.record()
        for _ in range(args.steps):
            metrics = step_resident()
        e1.record()
        barrier()
        records = _lib.profile_stop()
        profiled_ms_per_step = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the public API from pinned host memory ----
        h_frames, h_ids = synthetic_inputs(rank, device, pinned=True)
        chunk = 250  # videos per H2D chunk; copies run on a side stream and overlap the previous chunk's encode
        copy_stream = torch.cuda.Stream(device)
        n_chunks = VIDEOS_PER_GPU // chunk
        assert n_chunks % 2 == 0

        def time_e2e(h_video, encode):
            """Steady-state streaming of `h_video` (pinned host frames, any dtype) through `encode`: chunk i+1 is copied
            while chunk i is encoded, and the first chunk of the NEXT step starts streaming in while this step finishes
            (text encode, ranks, metrics, D2H) -- what a prefetching data loader does.  Every step issues exactly
            n_chunks H2D chunk copies."""
            bufs = [torch.empty((chunk, *h_video.shape[1:]), device=device, dtype=h_video.dtype) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            freed = [torch.cuda.Event() for _ in range(2)]

            def issue_copy(i):
                b = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[b])
                    bufs[b].copy_(h_video[i * chunk:(i + 1) * chunk], non_blocking=True)
                    ready[b].record(copy_stream)

            def step():
                main = torch.cuda.current_stream(device)
                outs = []
                for i in range(n_chunks):
                    if i + 1 < n_chunks:
                        issue_copy(i + 1)
                    b = i % 2
                    main.wait_event(ready[b])
                    outs.append(encode(bufs[b]))
                    freed[b].record(main)
                issue_copy(0)  # next step's first chunk
                d_ids = h_ids.to(device, non_blocking=True)
                t = encoder.encode_text({"input_ids": d_ids})
                ranks = retrieval_ranks(t, torch.cat(outs), group=group, totals=(n_total, n_total))
                m = metrics_from_ranks(ranks, n_total)
                return {k: x.cpu() for k, x in m.items()}  # D2H read of the step's result

            for b in range(2):
                freed[b].record(torch.cuda.current_stream(device))
            issue_copy(0)
            for _ in range(2):
                step()
            barrier()
            e0.record()
            for _ in range(args.steps):
                step()
            torch.cuda.current_stream(device).wait_event(ready[0])  # the last prefetch counts towards the timed region
            e1.record()
            barrier()
            t_ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            return t_ms.item() / args.steps

        e2e_ms = time_e2e(h_frames, encoder.encode_video)
        h2d_bytes = int(h_frames.numel() * 4 + h_ids.numel() * 4)
        # the reference's real input: decoded uint8 frames (here 240 x 320 RGB); its DataLoader workers run the eval
        # transform on the CPU (clip_video_text_encoder.py:124-133), here it is fused into the patch gather on the GPU
        gu = torch.Generator().manual_seed(4321 + rank)
        h_u8 = torch.empty(VIDEOS_PER_GPU, FRAMES, 240, 320, 3, dtype=torch.uint8, pin_memory=True)
        h_u8.random_(0, 256, generator=gu)
        e2e_u8_ms = time_e2e(h_u8, encoder.encode_video_uint8)
        e2e_uint8 = {"value": n_total / (e2e_u8_ms * 1e-3), "unit": "videos/s", "ms_per_step": e2e_u8_ms,
                     "h2d_bytes_per_step": int(h_u8.numel() + h_ids.numel() * 4), "d2h_bytes_per_step": 3 * 4 + 8,
                     "input": "uint8 decoded frames 240x320x3 from pinned host memory; eval transform (bicubic resize, "
                              "crop, normalise) fused into the patch gather (fc_encode_video_uint8)"}
        del h_u8

        # ---- outside every timed region: multi-GPU integer-rank equality (BASELINE.md section 4 gate) ----
        ranks_equal = None
        if world > 1:
            ranks_equal = rank_equality_check(encoder, frames, ids, n_total, world, rank, group, device)
        # ---- BASELINE configs[3] (the north_star target): 100k videos x 8 frames + 100k captions, STRONG scaling ----
        webvid = None
        if args.webvid_videos > 0:
            del frames, h_frames
            torch.cuda.empty_cache()
            webvid = webvid_leg(encoder, device, rank, world, group, args.webvid_videos, barrier)
        torch.cuda.empty_cache()
    train = None
    if world == 1 and args.train_videos > 0:  # row f3 (outside inference_mode: the trainer writes parameters in place)
        train = train_leg(encoder, device, args.train_videos)
    torch.cuda.empty_cache()  # the engines cudaMalloc their arenas: hand torch's cached blocks (training leg) back first
    geometries = geometry_leg(device, args.geometry_videos) if world == 1 and args.geometry_videos > 0 else None

    if rank == 0:
        peak, sustained, hbm, src = measured_peaks()
        gemm = [r for r in records if r["kind"] == 0]
        g_ms = sum(r["ms"] for r in gemm)
        g_flops = sum(r["flops"] for r in gemm)
        total_ms = sum(r["ms"] for r in records)
        achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
        by_kind = {}
        for r in records:
            name = {0: "gemm_tcgen05", 1: "attention", 2: "layernorm", 3: "other"}[r["kind"]]
            by_kind[name] = by_kind.get(name, 0.0) + r["ms"]
        shapes = sorted(gemm, key=lambda r: -r["ms"])
        step_flops = VIDEOS_PER_GPU * FRAMES * FLOP_PER_FRAME + CAPTIONS_PER_GPU * FLOP_PER_CAPTION \
            + 2.0 * n_total * VIDEOS_PER_GPU * 512
        traffic = None
        traffic_path = os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")  # the fc1 launch at the pass size run here
        if not os.path.exists(traffic_path):
            traffic_path = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
        if os.path.exists(traffic_path):  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from ncu --set full
            with open(traffic_path) as f:
                traffic = json.load(f)
        line = {
            "metric": "videos/sec (ViT-B/16 encode+sim+rank)", "value": n_total / (ms_per_step * 1e-3),
            "unit": "videos/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init CLIP ViT-B/16, N(0,1) frames, random token ids)",
            "config": {"workload": WORKLOAD, "videos_per_gpu": VIDEOS_PER_GPU, "frames_per_video": FRAMES,
                       "captions_per_gpu": CAPTIONS_PER_GPU, "gallery": n_total, "similarity": "split-bf16 x3",
                       "l2": "inputs (2.4 GB/GPU) exceed L2; no explicit flush", "parallelism": f"dp{world}",
                       "frames_per_pass": encoder.model.max_frames_per_pass},
            "queries_per_sec": n_total / (ms_per_step * 1e-3),
            "e2e_roofline_frac": step_flops / (ms_per_step * 1e-3) / (peak * 1e12),
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "videos/s",
                    "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 3 * 4 + 8, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "profiled_ms_per_step": profiled_ms_per_step,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_kernel": traffic["kernel"] if traffic else None,
                         "traffic_algorithmic_bytes": traffic["algorithmic_bytes_per_launch"] if traffic else None,
                         "peak_source": src,
                         "frac_of_sustained": achieved / sustained, "kernel": "gemm_bf16_tn_kernel (all shapes)",
                         "kernel_share_of_step": g_ms / total_ms if total_ms else None,
                         "ms_by_kernel_class": by_kind,
                         "top_shapes": [{"epi": r["tag"], "N": r["n"], "K": r["k"], "launches": r["launches"],
                                         "ms": round(r["ms"], 3),
                                         "tflops": round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 1)}
                                        for r in shapes[:8]]},
            "metrics": {k: float(v) for k, v in metrics.items()},
        }
        line.setdefault("extra", {})["e2e_uint8"] = e2e_uint8
        if ranks_equal is not None:
            line["ranks_equal_single_gpu"] = ranks_equal["equal"] and ranks_equal["equal_uneven"]
            line["ranks_equal_detail"] = ranks_equal
        if webvid is not None:
            total_flops = webvid["videos"] * (WEBVID_FRAMES * FLOP_PER_FRAME + FLOP_PER_CAPTION) \
                + 2.0 * 2 * webvid["videos"] ** 2 * 3 * 512
            webvid["algorithmic_tflop"] = total_flops / 1e12
            webvid["roofline_frac"] = total_flops / webvid["seconds"] / (world * peak * 1e12)
            webvid["roofline_frac_of_sustained"] = total_flops / webvid["seconds"] / (world * sustained * 1e12)
            ref_path = os.path.join(ROOT, "profiles", "r2_webvid_1gpu.json")
            webvid["kernel_sources_sha256"] = kernel_sources_sha256()
            if args.save_webvid_record and world == 1:
                with open(args.save_webvid_record, "w") as f:
                    json.dump(webvid, f, indent=1)
            if os.path.exists(ref_path):  # the N=1 run of this leg (same workload, same inputs), kept in profiles/
                with open(ref_path) as f:
                    one = json.load(f)
                if one.get("videos") == webvid["videos"]:
                    webvid["strong_scaling_efficiency"] = webvid["videos_per_s"] / (world * one["videos_per_s"])
                    # video i / caption i do not depend on the number of ranks, so R@k / MdR must be IDENTICAL -- as long as
                    # the record was made by the same kernels (any change of rounding order moves a few of 10^10 score
                    # comparisons): a record from other kernel sources is reported as stale, not as a mismatch
                    same_kernels = one.get("kernel_sources_sha256") == webvid["kernel_sources_sha256"]
                    webvid["metrics_equal_single_gpu"] = (webvid["metrics"] == one["metrics"]) if same_kernels else None
                    webvid["single_gpu_reference"] = {"videos_per_s": one["videos_per_s"], "metrics": one["metrics"],
                                                      "file": "profiles/r2_webvid_1gpu.json",
                                                      "same_kernel_sources": same_kernels}
            line.setdefault("extra", {})["webvid"] = webvid
        if train is not None:
            train["roofline_frac"] = train["achieved_tflops"] / peak
            line.setdefault("extra", {})["train_step"] = train
        if geometries is not None:
            for rec in geometries:
                rec["roofline_frac"] = rec["achieved_tflops"] / peak
            line.setdefault("extra", {})["geometries"] = geometries
        line["cpu_baseline"] = cpu_baseline(sample_videos=args.cpu_sample)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_sources_sha256() -> str:
    """Fingerprint of the CUDA sources the library was built from (the built .so is not byte-reproducible)."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for path in sorted(glob.glob(os.path.join(ROOT, "fitclip_b200", "csrc", "*.cu*"))):
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode() + b"\0" + f.read())
    return h.hexdigest()


def cpu_baseline(sample_videos: int = 128, steps: int = 1) -> dict:
    """The oracle (the reference's path restated: fp32 torch on the host cores) on a bounded sample of the workload:
    `sample_videos` videos x 4 frames + as many captions in batches of 32 (aligner/data/video_data_module.py:32),
    plus the full 1000 x 1000 similarity + argsort rank + metrics; encode time is extrapolated linearly to 1000."""
    import torch

    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = oracle.CLIP(**oracle.clip_ref.VIT_B_16).float().eval()
    model.load_state_dict(synthetic_weights(0))  # the same weights as the CUDA arm
    ref = oracle.RefClipVideoTextEncoder(model)
    g = torch.Generator().manual_seed(1234)
    video = torch.randn(sample_videos, FRAMES, 3, 224, 224, generator=g)
    ids = synthetic_tokens(sample_videos, 4321)
    with torch.inference_mode():
        ref(video[:2], {"input_ids": ids[:2]})  # warm-up
        t0 = time.perf_counter()
        for _ in range(steps):
            outs = [ref(video[i:i + 32], {"input_ids": ids[i:i + 32]}) for i in range(0, sample_videos, 32)]
        t_enc = (time.perf_counter() - t0) / steps
        # full-size similarity + argsort rank + metrics on the oracle's OWN embeddings: the sample's real embeddings
        # cycled to 1000 rows with a 1e-3 jitter (so rows differ), re-normalised
        ev, et = torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
        reps = -(-VIDEOS_PER_GPU // sample_videos)
        tv = torch.nn.functional.normalize(ev.repeat(reps, 1)[:VIDEOS_PER_GPU]
                                           + 1e-3 * torch.randn(VIDEOS_PER_GPU, ev.shape[1], generator=g), dim=-1)
        tt = torch.nn.functional.normalize(et.repeat(reps, 1)[:CAPTIONS_PER_GPU]
                                           + 1e-3 * torch.randn(CAPTIONS_PER_GPU, et.shape[1], generator=g), dim=-1)
        t0 = time.perf_counter()
        oracle.ref_retrieval_metrics(tt @ tv.T)
        t_rank = time.perf_counter() - t0
    total = t_enc * (VIDEOS_PER_GPU / sample_videos) + t_rank
    return {"value": VIDEOS_PER_GPU / total, "unit": "videos/s", "cores": cores, "kind": "port", "extrapolated": True,
            "sample": f"{sample_videos} videos x {FRAMES} frames + {sample_videos} captions encoded in {t_enc:.2f} s "
                      f"(encode time extrapolated x{VIDEOS_PER_GPU / sample_videos:.1f} to 1000) + full 1000x1000 "
                      f"sim+rank on the oracle's own embeddings {t_rank:.3f} s"}


def run_reference(args) -> None:
    """Reference arm: the reference's own CPU implementation of the path.  The reference package cannot be imported
    (clip / pytorch_lightning / torchmetrics / overrides absent, no network), so this times the oracle port on all
    host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    base = cpu_baseline(sample_videos=args.cpu_sample, steps=1)
    for _ in range(steps - 1):
        if time.perf_counter() - t0 > 150:
            break
        base = cpu_baseline(sample_videos=args.cpu_sample, steps=1)
    v = base["value"]
    print(json.dumps({
        "impl": "reference", "extrapolated": True,
        "metric": "videos/sec (ViT-B/16 encode+sim+rank)", "value": v, "unit": "videos/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * VIDEOS_PER_GPU / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference path; each step is a bounded sample (128 videos + 128 "
                                                "captions) whose encode time is extrapolated linearly to the 1000-video "
                                                "workload: value and ms_per_step are EXTRAPOLATED, not a full run"},
        "cpu_baseline": base, "gpu_launches": 0,
        "e2e": {"value": v, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--webvid-videos", type=int, default=100_000,
                    help="gallery size of the BASELINE configs[3] leg (100k videos x 8 frames, split over the ranks; "
                         "reported under extra.webvid, outside the K timed steps); 0 skips it")
    ap.add_argument("--save-webvid-record", default=None,
                    help="N=1 only: also write the 100k-video leg's record to this path (profiles/r2_webvid_1gpu.json is "
                         "the single-GPU reference the multi-GPU runs compare their metrics and time with)")
    ap.add_argument("--frames-per-pass", type=int, default=None,
                    help="frames per internal encoder pass (default: the library's token budget, ~2000 frames of ViT-B/16)")
    ap.add_argument("--train-videos", type=int, default=512,
                    help="videos per step of the teacher-student TRAINING leg (row f3; extra.train_step, N = 1 only); 0 skips it")
    ap.add_argument("--geometry-videos", type=int, default=128,
                    help="N=1 only: videos per step of the other-geometry legs (ViT-L/14, SLIP-layout ViT-B/16); 0 skips them")
    ap.add_argument("--cpu-sample", type=int, default=128,
                    help="videos in the bounded CPU-baseline sample (128 videos x 4 frames + 128 captions: 10-30 s of CPU work)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
