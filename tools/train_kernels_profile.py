"""One launch of every training-step kernel at the step's shapes (256 frames = one eighth of the 2048-frame step), for
    ncu --set full --clock-control none --import-source on -k regex:'attention_bwd|gemm_bf16_tn|ln_bwd|colsum|quickgelu' \
        -o gpurun_out/r1_train_kernels python tools/train_kernels_profile.py
(summarised with tools/ncu_summary.py into profiles/r1_ncu_train_kernels_summary.csv)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import ops, train_ops as T  # noqa: E402

dev = torch.device("cuda:0")
F, L, W, H = 256, 197, 768, 12
M = F * L
bf = torch.bfloat16
x = torch.randn(M, W, device=dev).to(bf)
dy3 = torch.randn(M, 3 * W, device=dev).to(bf)
du = torch.randn(M, 4 * W, device=dev).to(bf)
u = torch.randn(M, 4 * W, device=dev).to(bf)
w_fc = (torch.randn(4 * W, W, device=dev) / W ** 0.5).to(bf)
g_fc = torch.zeros(4 * W, W, device=dev)
g_qkv = torch.zeros(3 * W, W, device=dev)
gamma = torch.ones(W, device=dev)
dg, db, cs = torch.zeros(W, device=dev), torch.zeros(W, device=dev), torch.zeros(4 * W, device=dev)
zero_bias = torch.zeros(W, device=dev)
qkv = torch.randn(M, 3 * W, device=dev).to(bf)
out = ops.attention_bf16(qkv, F, L, H, False)
torch.cuda.synchronize()

T.attention_bwd(qkv, out, x, F, L, H, False)          # attention backward (tcgen05), 256 x 197 x 12
T.wgrad_tn(du, x, g_fc)                               # wgrad fc1: (3072 x 768) += dY^T X, K = 50432 token rows, split-K
T.wgrad_tn(dy3, x, g_qkv)                             # wgrad qkv: (2304 x 768)
T.gemm_nt(du, w_fc, zero_bias)                        # dgrad fc1: dX = dY W with W read in place (MN-major B)
T.layernorm_bwd(x, x, gamma, dg, db, add=x)           # LayerNorm backward + residual gradient
T.colsum(du, cs)                                      # bias gradient, 3072 columns
T.quickgelu(u)                                        # QuickGELU forward
T.quickgelu_bwd(u, du, g_out=torch.empty_like(u))     # QuickGELU backward (+ activation)
torch.cuda.synchronize()
