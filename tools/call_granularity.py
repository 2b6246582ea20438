"""Throughput at the REFERENCE's call granularity: the reference evaluates with batch_size 32 (aligner/data/
video_data_module.py:32), i.e. one encode_video call per 32 videos x 4 frames = 128 frames and one encode_text call per 32
captions -- against the bench's one-call-per-1000-videos step.  Shows what per-launch host work (tensor-map encodes,
~75 launches per pass) costs when a pass is 4x smaller than the engine's 500-frame passes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import B200ClipVideoTextEncoder, metrics_from_ranks, retrieval_ranks  # noqa: E402

dev = torch.device("cuda:0")
enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0).state_dict(), num_frames=4).to(dev)
gd = torch.Generator(device=dev).manual_seed(1234)
frames = torch.randn(1000, 4, 3, 224, 224, device=dev, generator=gd)
ids = oracle.tokenize_synthetic(1000, 77, seed=4321).to(dev)


def step(bs):
    v = torch.cat([enc.encode_video(frames[i:i + bs]) for i in range(0, 1000, bs)])
    t = torch.cat([enc.encode_text({"input_ids": ids[i:i + bs]}) for i in range(0, 1000, bs)])
    return metrics_from_ranks(retrieval_ranks(t, v), 1000)


out = {}
with torch.inference_mode():
    for bs in (1000, 125, 32, 8):
        for _ in range(2):
            step(bs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step(bs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[f"batch_{bs}"] = {"ms_per_1000_videos": ms, "videos_per_s": 1e6 / ms}
print(json.dumps(out, indent=1))
