"""Sustained throughput of the fused attention kernel at the bench shape (256 frames x 12 heads x 197 tokens)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import _lib  # noqa: E402

# optional argument: library variant ("timing": make TIMING=1, or any `make VARIANT=name EXTRA=...` build)
VARIANT = sys.argv[1] if len(sys.argv) > 1 else ""
TIMING = VARIANT == "timing"
if VARIANT:
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libfitclip_b200.so", f"libfitclip_b200_{VARIANT}.so")
from fitclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for seqs, L, heads, causal in [(256, 197, 12, False), (984, 77, 8, True)]:
    qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()
    for _ in range(5):
        ops.attention_bf16(qkv, seqs, L, heads, causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for _ in range(reps):
        ops.attention_bf16(qkv, seqs, L, heads, causal)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    flops = 4.0 * seqs * heads * L * L * 64
    exps = seqs * heads * L * L
    print(f"[{VARIANT or 'default'}] attention seqs={seqs} L={L} heads={heads} causal={causal}: {us:7.1f} us  {flops / us / 1e6:6.1f} TFLOP/s  "
          f"{exps / us / 1e3:6.2f} Gexp/s")
    if TIMING and not causal:
        import ctypes as C
        lib = _lib.load()
        lib.fc_debug_att_timing.argtypes = [C.c_void_p, C.c_int]
        buf = (C.c_ulonglong * 8)()
        lib.fc_debug_att_timing(None, 1)
        ops.attention_bf16(qkv, seqs, L, heads, causal)
        lib.fc_debug_att_timing(buf, 1)
        t = [int(v) for v in buf]
        n = max(t[7], 1)
        names = ["loop", "wait S main", "pass 1", "wait S tail", "pass 2", "wait O", "epilogue"]
        print("  softmax warp 0, cycles per item: " + ", ".join(f"{nm} {t[i] / n:.0f}" for i, nm in enumerate(names)) +
              f"  ({n} items)")
