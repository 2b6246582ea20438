"""Two launches of the fused attention kernel at the bench shape, for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
qkv = torch.randn(256 * 197, 3 * 12 * 64, device=dev).bfloat16()
for _ in range(2):
    ops.attention_bf16(qkv, 256, 197, 12, False)
torch.cuda.synchronize()
