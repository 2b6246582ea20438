"""Achieved HBM bandwidth of the memory-bound kernels of the path against the measured copy peak
(MEASURED_PEAKS.json `hbm_gbs`).  Each case uses buffers larger than the 126 MB L2, algorithmic bytes (SURVEY.md 8d /
DESIGN.md section 3) divided by the CUDA-event time of back-to-back launches.  Prints one JSON line per kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fitclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
peak = 6550.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def report(name, nbytes, seconds, note):
    gbs = nbytes / seconds / 1e9
    print(json.dumps({"kernel": name, "algorithmic_bytes": int(nbytes), "us": round(seconds * 1e6, 1),
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3), "note": note}))


# K14 WiSE lerp over the ViT-B/16 parameter count: 2 fp32 reads + 1 fp32 write per parameter
n = 149_620_737
a, b, o = (torch.randn(n, device=dev) for _ in range(3))
report("wise_lerp_kernel", 12.0 * n, timed(lambda: ops.wise_lerp(a, b, 0.4, out=o)), "149.6 M fp32 parameters, w = 0.4")
del a, b, o

# K2 LayerNorm (ln_pre): bf16 in + bf16 out, 512 frames x 197 tokens x 768
rows, D = 512 * 197, 768
x = torch.randn(rows, D, device=dev).bfloat16()
g, be = torch.ones(D, device=dev), torch.zeros(D, device=dev)
y = torch.empty_like(x)
report("layernorm_bf16_kernel", 4.0 * rows * D, timed(lambda: ops.layernorm_bf16(x, g, be, out=y)),
       "100864 rows x 768 (one 512-frame pass); 155 MB in + 155 MB out")
del x, y
rows = 2000 * 197  # the pass size the engine actually runs (token budget: ~2000 frames of ViT-B/16)
x = torch.randn(rows, D, device=dev).bfloat16()
y = torch.empty_like(x)
report("layernorm_bf16_kernel (2000-frame pass)", 4.0 * rows * D, timed(lambda: ops.layernorm_bf16(x, g, be, out=y)),
       "394000 rows x 768; 605 MB in + 605 MB out")
del x, y

# K8 pool + normalise: fp32 (B*T, 512) -> (B, 512)
B, T = 200_000, 8
f = torch.randn(B * T, 512, device=dev)
report("pool_normalize_kernel", 4.0 * (B * T + B) * 512, timed(lambda: ops.pool_normalize(f, T)),
       "200k videos x 8 frames x 512")
del f

# f1 pre-processing: uint8 (n, 360, 640, 3) -> bf16 (n, 3, 224, 224)
nfr = 512
raw = torch.randint(0, 256, (nfr, 360, 640, 3), dtype=torch.uint8, device=dev)
mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
# only the centre-cropped window is read: 360 x 360 source pixels per frame
report("preprocess_kernel", nfr * (360.0 * 360 * 3 + 3 * 224 * 224 * 2),
       timed(lambda: ops.preprocess_frames(raw, 224, mean, std, torch.bfloat16)),
       "512 frames 360x640 -> 224x224 bf16 (reads the 360x360 crop window)")
del raw

# K13 rank from a materialised score matrix: 4 bytes per score
nt = nv = 16384
s = torch.randn(nt, nv, device=dev)
tgt = torch.arange(nt, device=dev, dtype=torch.int32)
report("rank_from_scores_kernel", 4.0 * nt * nv, timed(lambda: ops.rank_from_scores(s, tgt)), "16384 x 16384 fp32 scores (1.07 GB)")
del s, tgt

# ---- training-step kernels (row f3)
from fitclip_b200 import train_ops as T  # noqa: E402

# AdamW over the ViT-B/16 parameter count: p, g, m, v read (16 B), p, m, v written (12 B), bf16 mirror written (2 B)
n = 149_620_736
p_, g_, m_, v_ = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v_.abs_()
pb = torch.empty(n, device=dev, dtype=torch.bfloat16)
report("adamw_kernel", 30.0 * n, timed(lambda: T.adamw_step(p_, g_, m_, v_, 3, lr=3e-6, p_bf16=pb)),
       "149.6 M parameters: fp32 p / g / m / v in, p / m / v + bf16 mirror out")
del p_, g_, m_, v_, pb

# LayerNorm backward with the residual gradient added: x, dy, add read (6 B), dx written (2 B) per element
rows, D = 512 * 197, 768
x = torch.randn(rows, D, device=dev).bfloat16()
dy = torch.randn(rows, D, device=dev).bfloat16()
add = torch.randn(rows, D, device=dev).bfloat16()
gam = torch.ones(D, device=dev)
dgam, dbet = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
dx = torch.empty_like(x)
report("ln_bwd_kernel", 8.0 * rows * D, timed(lambda: T.layernorm_bwd(x, dy, gam, dgam, dbet, add=add, out=dx)),
       "100864 rows x 768: x, dy, residual gradient in, dx out (+ dgamma / dbeta)")
