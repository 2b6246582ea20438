"""One short pass of the hot path for ncu: `videos` x 4 frames (one vision pass when <= ~500 videos), 984 captions (one
text pass), similarity+rank.
Usage: python tools/profile_step.py [passes [videos]]     (videos = 500 is the 2000-frame pass the bench runs)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import B200ClipVideoTextEncoder, metrics_from_ranks, retrieval_ranks  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0).state_dict()).to(dev)
videos = int(sys.argv[2]) if len(sys.argv) > 2 else 64
video = torch.randn(videos, 4, 3, 224, 224, device=dev)
ids = oracle.tokenize_synthetic(984, 77).to(dev)
with torch.inference_mode():
    for _ in range(passes):
        v = enc.encode_video(video)
        t = enc.encode_text({"input_ids": ids})
        ranks = retrieval_ranks(t[:64].contiguous(), v)
        m = metrics_from_ranks(ranks, 64)
torch.cuda.synchronize()
print({k: float(x) for k, x in m.items()})
