"""attention_tc_long_kernel (optionally a `make VARIANT=...` build: FITCLIP_VARIANT) against a torch fp32 reference at the
long image sequence lengths, plus its timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import _lib as _libmod  # noqa: E402

if os.environ.get("FITCLIP_VARIANT"):
    _libmod.LIB_PATH = _libmod.LIB_PATH.replace("libfitclip_b200.so", "libfitclip_b200_%s.so" % os.environ["FITCLIP_VARIANT"])
from fitclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
tag = os.environ.get("FITCLIP_VARIANT", "default")
for seqs, L, heads in [(3, 257, 16), (40, 257, 16), (2, 577, 4), (10, 577, 16), (3, 768, 2), (5, 256, 4), (6, 385, 3), (2, 209, 3),
                       (4, 320, 2)]:
    g = torch.Generator(device=dev).manual_seed(L + seqs)
    qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev, generator=g).bfloat16()
    out = ops.attention_bf16(qkv, seqs, L, heads, False)
    q, k, v = (qkv.float().view(seqs, L, 3, heads, 64)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).permute(0, 2, 1, 3).reshape(seqs * L, heads * 64)
    err = (out.float() - ref).abs().max().item()
    print(f"[{tag}] seqs={seqs} L={L} heads={heads}: max abs err {err:.3e}", flush=True)
    assert err < 3e-2, err
for seqs, L, heads in [(512, 257, 16), (512, 577, 16)]:
    qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()
    for _ in range(3):
        ops.attention_bf16(qkv, seqs, L, heads, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention_bf16(qkv, seqs, L, heads, False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"[{tag}] seqs={seqs} L={L} heads={heads}: {us:8.1f} us  {4.0 * seqs * heads * L * L * 64 / us / 1e6:6.1f} TFLOP/s")
