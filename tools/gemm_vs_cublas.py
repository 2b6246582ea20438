"""Sustained-throughput comparison of the tcgen05 GEMM with cuBLAS (torch.matmul) on the four ViT-B/16 layer shapes.
Each case runs back to back for ~1.5 s so the power cap / clocks settle (that is the regime the encoder runs in)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import _lib as _libmod  # noqa: E402

if os.environ.get("FITCLIP_VARIANT"):  # A/B runs against a `make VARIANT=name` build of the library
    _libmod.LIB_PATH = _libmod.LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_%s.so" % os.environ["FITCLIP_VARIANT"])
from fitclip_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
M = 50432
shapes = [("qkv", 2304, 768, _lib.EPI_BIAS), ("out", 768, 768, _lib.EPI_BIAS_RESID), ("fc1", 3072, 768, _lib.EPI_BIAS_QGELU),
          ("fc2", 768, 3072, _lib.EPI_BIAS_RESID)]


def sustained(fn, flops, seconds=1.5):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return flops * n / (ms * 1e-3) / 1e12, ms / n * 1e3


for name, N, K, epi in shapes:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    x = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * M * N * K
    resid = x if epi == _lib.EPI_BIAS_RESID else None
    ours = sustained(lambda: ops.gemm_bf16(a, b, bias, resid=resid, epilogue=epi, out=out), flops)
    bt = b.t()
    cublas = sustained(lambda: torch.matmul(a, bt, out=out), flops)
    print(f"{name}: M={M} N={N} K={K}  ours {ours[0]:7.1f} TF/s ({ours[1]:6.1f} us)   cuBLAS plain matmul {cublas[0]:7.1f} TF/s "
          f"({cublas[1]:6.1f} us)")
