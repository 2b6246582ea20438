"""A tiny pass through every eval-path kernel family (encoder with the 197-token and 77-token tcgen05 attention, the
key-block attention at 257 tokens, fused uint8 pre-processing, similarity + rank count) for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import B200Clip, B200ClipVideoTextEncoder, metrics_from_ranks, ops, retrieval_ranks  # noqa: E402

dev = torch.device("cuda:0")
sd = oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1).state_dict()
enc = B200ClipVideoTextEncoder(B200Clip(sd, max_frames_per_pass=8, max_texts_per_pass=8), num_frames=2).to(dev)
video = torch.randn(3, 2, 3, 224, 224, device=dev)
ids = oracle.tokenize_synthetic(5, (4, 77), seed=1).to(dev)
with torch.inference_mode():
    v = enc.encode_video(video)
    t = enc.encode_text({"input_ids": ids})
    raw = torch.randint(0, 256, (2, 2, 120, 160, 3), dtype=torch.uint8, device=dev)
    u = enc.encode_video_uint8(raw)
    ranks = retrieval_ranks(t[:3].contiguous(), v)
    m = metrics_from_ranks(ranks, 3)
    qkv = torch.randn(2 * 257, 3 * 2 * 64, device=dev).bfloat16()
    o = ops.attention_bf16(qkv, 2, 257, 2, False)
    qkv = torch.randn(3 * 50, 3 * 2 * 64, device=dev).bfloat16()
    o2 = ops.attention_bf16(qkv, 3, 50, 2, False)
torch.cuda.synchronize()
print("sanitizer smoke ok", v.shape, t.shape, u.shape, ranks.tolist(), {k: float(x) for k, x in m.items()}, o.shape, o2.shape)
