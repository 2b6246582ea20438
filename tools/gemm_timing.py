"""Where the GEMM's cycles go, per warp role (diagnostics build: `make -C fitclip_b200/csrc TIMING=1`).
For each ViT-B/16 layer shape: cycles per tile of the MMA thread, split into waiting for a drained accumulator
(epilogue-bound), waiting for TMA bytes (feed-bound) and issuing; the producer's wait for a free stage; the epilogue's
wait for accumulators."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import _lib  # noqa: E402

_lib.LIB_PATH = _lib.LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_timing.so")
from fitclip_b200 import ops  # noqa: E402

lib = _lib.load()
lib.fc_debug_gemm_timing.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50432
shapes = [("qkv", 2304, 768, _lib.EPI_BIAS), ("out", 768, 768, _lib.EPI_BIAS_RESID), ("fc1", 3072, 768, _lib.EPI_BIAS_QGELU),
          ("fc2", 768, 3072, _lib.EPI_BIAS_RESID), ("sim", 1000, 1536, _lib.EPI_F32)]
buf = (C.c_ulonglong * 8)()
for name, N, K, epi in shapes:
    m = 1000 if name == "sim" else M
    a = (torch.randn(m, K, device=dev) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    x = torch.randn(m, N, device=dev).bfloat16()
    out = torch.empty(m, N, device=dev, dtype=torch.float32 if epi == _lib.EPI_F32 else torch.bfloat16)
    resid = x if epi == _lib.EPI_BIAS_RESID else None
    kw = dict(epilogue=epi, out=out)
    if epi == _lib.EPI_F32:
        fn = lambda: ops.gemm_bf16(a, b, **kw)
    else:
        fn = lambda: ops.gemm_bf16(a, b, bias, resid=resid, **kw)
    for _ in range(5):
        fn()
    lib.fc_debug_gemm_timing(None, 1)
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    lib.fc_debug_gemm_timing(buf, 1)
    t = [int(v) for v in buf]
    tiles, mma_threads = max(t[6], 1), max(t[7], 1)
    nk = (K + 63) // 64
    print(f"{name}: M={m} N={N} K={K}  {us:7.1f} us  {2.0 * m * N * K / us / 1e6:7.1f} TF/s | per pair-tile (cycles): "
          f"mma loop {t[0] / tiles:8.0f} = wait-acc {t[1] / tiles:7.0f} + wait-tma {t[2] / tiles:7.0f} "
          f"({t[2] / tiles / nk:5.0f}/kblock) + issue {(t[0] - t[1] - t[2]) / tiles:7.0f} ({(t[0] - t[1] - t[2]) / tiles / nk:5.0f}/kblock)"
          f" | producer wait-stage {t[3] / (2 * tiles):7.0f} | epilogue loop {t[4] / tiles:8.0f}, wait-acc {t[5] / tiles:7.0f}"
          f" | tiles/launch {tiles / reps:.0f} over {mma_threads / reps:.0f} pairs")
