"""Per-kernel times of one 500-frame vision pass (the library's CUDA-event profiler, fc_profile_start/stop), sorted by
time: the place to look at the small kernels (patch-embed GEMM = kind 0 tag 3, head = kind 3 tag 3, im2col = kind 3 tag 0,
ln_pre = kind 2, pool = kind 3 tag 4).  FITCLIP_VARIANT=name runs a `make VARIANT=name` build for A/B comparisons."""
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
