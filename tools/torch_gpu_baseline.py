"""The PyTorch 2.11 LIBRARY path on the same B200, beside our numbers (SURVEY.md 2.2's bar; a diagnostic, never a
product path): the oracle's CLIP cast to bf16 on the GPU -- cuBLASLt GEMMs, nn.MultiheadAttention's fused SDPA / flash
kernels, torch LayerNorm -- for the bench's cfg2 step (1000 videos x 4 frames + 1000 captions -> similarity -> ranks),
eager and (optionally) torch.compile, plus per-kernel attention: F.scaled_dot_product_attention against
attention_tc_kernel at the two CLIP shapes.

    python tools/torch_gpu_baseline.py [--compile] [--steps 5] > profiles/r2_torch_gpu_baseline.json
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import B200ClipVideoTextEncoder, metrics_from_ranks, ops, retrieval_ranks  # noqa: E402

dev = torch.device("cuda:0")
FLOP_PER_FRAME, FLOP_PER_CAPTION = 35_126_906_880, 5_959_540_736


def timed(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                              "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        return out
    except OSError:
        return None


def attention_kernels(out):
    """F.scaled_dot_product_attention (flash / cuDNN backends as torch selects) vs attention_tc_kernel."""
    rows = []
    for name, seqs, L, heads, causal in (("image 4000x12 heads, 197 tok", 4000, 197, 12, False),
                                         ("image 500x12 heads, 197 tok (one pass)", 500, 197, 12, False),
                                         ("text 1000x8 heads, 77 tok causal", 1000, 77, 8, True)):
        qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()
        q, k, v = (qkv.view(seqs, L, 3, heads, 64)[:, :, i].permute(0, 2, 1, 3).contiguous() for i in range(3))
        flops = 4.0 * seqs * heads * L * L * 64
        ms_ours = timed(lambda: ops.attention_bf16(qkv, seqs, L, heads, causal), 50)
        ms_sdpa = timed(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=causal), 50)
        ref = F.scaled_dot_product_attention(q, k, v, is_causal=causal).permute(0, 2, 1, 3).reshape(seqs * L, heads * 64)
        got = ops.attention_bf16(qkv, seqs, L, heads, causal)
        rows.append({"shape": name, "ours_ms": ms_ours, "sdpa_ms": ms_sdpa, "ours_tflops": flops / ms_ours / 1e9,
                     "sdpa_tflops": flops / ms_sdpa / 1e9, "speedup_vs_sdpa": ms_sdpa / ms_ours,
                     "max_abs_diff_vs_sdpa": (got.float() - ref.float()).abs().max().item(),
                     "note": "SDPA timed on pre-split contiguous (B,H,L,64) q/k/v: the split/permute copies it needs in "
                             "the real model are NOT charged to it"})
    out["attention"] = rows


def step_baseline(out, steps, do_compile):
    model = oracle.clip_vit_b_16(seed=0)
    sd = model.state_dict()
    ours = B200ClipVideoTextEncoder(sd, num_frames=4).to(dev)
    model = model.to(dev).bfloat16()
    for mod in model.modules():  # like clip.model.convert_weights: LayerNorms keep fp32 parameters (their forward casts)
        if isinstance(mod, torch.nn.LayerNorm):
            mod.float()
    lib = oracle.RefClipVideoTextEncoder(model)
    gd = torch.Generator(device=dev).manual_seed(1234)
    frames = torch.randn(1000, 4, 3, 224, 224, device=dev, generator=gd)
    ids = oracle.tokenize_synthetic(1000, 77, seed=4321).to(dev)
    frames_bf16 = frames.bfloat16()
    flops = 4000 * FLOP_PER_FRAME + 1000 * FLOP_PER_CAPTION + 2.0 * 1000 * 1000 * 512

    def lib_step(enc):
        # the reference's call granularity would be 32 videos per call; the library path gets the friendlier 250
        v = torch.cat([enc.encode_video(frames_bf16[i:i + 250]) for i in range(0, 1000, 250)]).float()
        t = enc.encode_text({"input_ids": ids}).float()
        scores = t @ v.T
        ranks = (scores > scores.diagonal().unsqueeze(1)).sum(dim=1)
        return (ranks < 1).float().mean(), (ranks < 5).float().mean(), (ranks < 10).float().mean(), ranks.median() + 1

    def our_step():
        v = ours.encode_video(frames)
        t = ours.encode_text({"input_ids": ids})
        return metrics_from_ranks(retrieval_ranks(t, v), 1000)

    with torch.inference_mode():
        ms_ours = timed(our_step, steps)
        c0 = clocks()
        ms_eager = timed(lambda: lib_step(lib), steps)
        c1 = clocks()
        rec = {"workload": "cfg2: 1000 videos x 4 frames + 1000 captions, ViT-B/16, similarity + ranks",
               "ours_ms": ms_ours, "torch_bf16_eager_ms": ms_eager, "ours_videos_per_s": 1e6 / ms_ours,
               "torch_bf16_eager_videos_per_s": 1e6 / ms_eager, "speedup_vs_torch_eager": ms_eager / ms_ours,
               "ours_tflops": flops / ms_ours / 1e9, "torch_eager_tflops": flops / ms_eager / 1e9,
               "clocks_after_ours": c0, "clocks_after_torch": c1,
               "note": "library path: oracle CLIP .bfloat16() on cuda (cuBLASLt, nn.MultiheadAttention fast path / SDPA, "
                       "torch LayerNorm), bf16 inputs pre-cast (not charged), 250-video calls"}
        # parity of the library path itself, for context (bf16 weights AND activations: looser than ours)
        v_lib = lib.encode_video(frames_bf16[:8]).float()
        v_ours = ours.encode_video(frames[:8])
        rec["cos_ours_vs_torch_bf16"] = F.cosine_similarity(v_lib, v_ours).min().item()
        if do_compile:
            try:
                t0 = time.time()
                import copy
                cm = copy.deepcopy(lib.model)
                cm.visual = torch.compile(cm.visual)          # the towers are what encode_image / encode_text call
                cm.transformer = torch.compile(cm.transformer)
                comp = oracle.RefClipVideoTextEncoder(cm)
                ms_comp = timed(lambda: lib_step(comp), steps, warmup=2)
                rec.update(torch_bf16_compile_ms=ms_comp, torch_bf16_compile_videos_per_s=1e6 / ms_comp,
                           speedup_vs_torch_compile=ms_comp / ms_ours, compile_seconds=time.time() - t0)
            except Exception as e:  # noqa: BLE001 -- a diagnostic tool: record why and go on
                rec["torch_compile_error"] = repr(e)[:300]
    out["step"] = rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--compile", action="store_true")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    with torch.inference_mode():
        attention_kernels(out)
    step_baseline(out, args.steps, args.compile)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
