"""Videos/s of the vision tower as a function of frames per pass (L2 residency of the activations vs GEMM wave fill).
Usage: python tools/pass_size_sweep.py [frames_per_pass ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import _lib as _libmod  # noqa: E402

if os.environ.get("FITCLIP_VARIANT"):  # A/B runs against a `make VARIANT=name` build of the library
    _libmod.LIB_PATH = _libmod.LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_%s.so" % os.environ["FITCLIP_VARIANT"])
import oracle  # noqa: E402
from fitclip_b200 import B200Clip, B200ClipVideoTextEncoder  # noqa: E402

dev = torch.device("cuda:0")
sizes = [int(a) for a in sys.argv[1:]] or [96, 192, 256, 384]
sd = oracle.clip_vit_b_16(seed=0).state_dict()
video = torch.randn(960, 4, 3, 224, 224, device=dev)
for f in sizes:
    enc = B200ClipVideoTextEncoder(B200Clip(sd, max_frames_per_pass=f)).to(dev)
    with torch.inference_mode():
        for _ in range(2):
            enc.encode_video(video)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 4
        for _ in range(reps):
            enc.encode_video(video)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"frames/pass {f:4d}: {ms:7.1f} ms per 960 videos -> {960 / ms * 1e3:7.1f} videos/s (vision tower only)")
    del enc
    torch.cuda.empty_cache()
