"""Row f4: the other CLIP geometries the reference ships configs for (config/encoder/clip_vit_b_32.yaml,
clip_vit_l_14.yaml, clip_vit_l_14_336px.yaml), random-init, on one GPU: retrieval-eval step of 256 videos x 4 frames +
256 captions -> videos/s and achieved TFLOP/s (dense algorithmic count), plus the attention kernel alone at each image
sequence length.  One JSON line per geometry."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import _lib as _libmod  # noqa: E402

if os.environ.get("FITCLIP_VARIANT"):  # A/B runs against a `make VARIANT=name` build of the library
    _libmod.LIB_PATH = _libmod.LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_%s.so" % os.environ["FITCLIP_VARIANT"])
import oracle  # noqa: E402
from fitclip_b200 import (B200ClipVideoTextEncoder, B200SlipVideoTextEncoder, metrics_from_ranks, ops,  # noqa: E402
                          retrieval_ranks)

dev = torch.device("cuda:0")
GEOM = {
    "clip_vit_b_32": dict(vision_patch_size=32),
    "clip_vit_b_16": dict(),
    "clip_vit_l_14": dict(embed_dim=768, vision_patch_size=14, vision_width=1024, vision_layers=24, transformer_width=768,
                          transformer_heads=12),
    "clip_vit_l_14_336px": dict(embed_dim=768, image_resolution=336, vision_patch_size=14, vision_width=1024,
                                vision_layers=24, transformer_width=768, transformer_heads=12),
    # SLIP layout (slip.py:595-600 / 618-623): timm ViT image tower (no ln_pre, exact GELU) + CLIP text tower
    "slip_vit_s_16": dict(slip=True, vision_width=384, vision_heads=12),
    "slip_vit_b_16": dict(slip=True),
    "slip_vit_l_16": dict(slip=True, vision_width=1024, vision_layers=24),
}


def tower_flops(tokens, width, layers, heads):
    per_layer = 2 * tokens * 12 * width * width + 4 * tokens * tokens * width  # heads * head_dim = width
    return layers * per_layer


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


names = sys.argv[1:] or list(GEOM)
for name in names:
    geom = dict(GEOM[name])
    if geom.pop("slip", False):
        cfg = {**oracle.clip_ref.VIT_B_16, **geom}
        from fitclip_b200 import B200SlipClip
        enc = B200SlipVideoTextEncoder(B200SlipClip(oracle.slip_clip_vit_b_16(seed=0, **geom).state_dict(),
                                                    vision_heads=geom.get("vision_heads")), num_frames=4).to(dev)
    else:
        cfg = {**oracle.clip_ref.VIT_B_16, **geom}
        enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **geom).state_dict(), num_frames=4).to(dev)
    res, patch = cfg["image_resolution"], cfg["vision_patch_size"]
    L = (res // patch) ** 2 + 1
    n = 256
    g = torch.Generator(device=dev).manual_seed(1)
    video = torch.randn(n, 4, 3, res, res, device=dev, generator=g)
    ids = oracle.tokenize_synthetic(n, 77, seed=2).to(dev)

    def step():
        v = enc.encode_video(video)
        t = enc.encode_text({"input_ids": ids})
        return metrics_from_ranks(retrieval_ranks(t, v), n)

    with torch.inference_mode():
        ms, _ = timed(step, 3)
        vw, heads = cfg["vision_width"], geom.get("vision_heads") or cfg["vision_width"] // 64
        frame = tower_flops(L, vw, cfg["vision_layers"], heads) + 2 * (L - 1) * vw * 3 * patch * patch + 2 * vw * cfg["embed_dim"]
        cap = tower_flops(77, cfg["transformer_width"], cfg["transformer_layers"], cfg["transformer_heads"]) \
            + 2 * cfg["transformer_width"] * cfg["embed_dim"]
        flops = n * 4 * frame + n * cap
        seqs = 512
        qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()  # narrow heads run in 64-wide slots
        att_ms, _ = timed(lambda: ops.attention_bf16(qkv, seqs, L, heads, False), 20)
        print(json.dumps({"geometry": name, "image_tokens": L, "videos": n, "frames_per_video": 4, "ms_per_step": round(ms, 2),
                          "videos_per_s": round(n / ms * 1e3, 1), "tflops": round(flops / ms / 1e9, 1),
                          "gflop_per_frame": round(frame / 1e9, 2),
                          "attention_only": {"seqs": seqs, "L": L, "heads": heads, "us": round(att_ms * 1e3, 1),
                                             "tflops": round(4.0 * seqs * heads * L * L * 64 / att_ms / 1e9, 1)}}), flush=True)
    del enc, video, qkv
    torch.cuda.empty_cache()
