"""Timing of the pooled-token head (LayerNorm + projection) through a 1-layer encoder is not separable, so this times
fc_encode paths indirectly: it calls the profiler around a full vision pass and prints the per-kernel-class times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import _lib  # noqa: E402

if os.environ.get("FITCLIP_VARIANT"):
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_%s.so" % os.environ["FITCLIP_VARIANT"])
from fitclip_b200 import B200ClipVideoTextEncoder  # noqa: E402

dev = torch.device("cuda:0")
enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0).state_dict()).to(dev)
video = torch.randn(125, 4, 3, 224, 224, device=dev)
with torch.inference_mode():
    for _ in range(3):
        enc.encode_video(video)
    torch.cuda.synchronize()
    _lib.profile_start(1 << 12)
    for _ in range(5):
        enc.encode_video(video)
    recs = _lib.profile_stop()
for r in sorted(recs, key=lambda r: -r["ms"]):
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("kind", "tag", "n", "k", "launches", "ms")})
