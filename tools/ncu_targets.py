"""Launches the round-2 kernels that have no ncu record yet, a few times each, for one `ncu --set full` capture
(tools/r2_ncu_capture2.sh): the long-sequence attention with 192-key blocks at the ViT-L/14 (257) and ViT-L/14@336 (577)
shapes, and the c_proj dgrad fused with QuickGELU's backward (EPI_QGELU_BWD) at the training step's shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import ops, train_ops  # noqa: E402

dev = torch.device("cuda:0")
for L in (257, 577):
    qkv = torch.randn(512 * L, 3 * 1024, device=dev).bfloat16()
    for _ in range(3):
        ops.attention_bf16(qkv, 512, L, 16, False)
M, N, K = 512 * 197, 3072, 768
dy = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(K, N, device=dev) * K ** -0.5).bfloat16()
u = (torch.randn(M, N, device=dev) * 2).bfloat16()
for _ in range(3):
    train_ops.gemm_nt(dy, w, None, qgelu_bwd_of=u)
torch.cuda.synchronize()
print("done")
