"""Times fc_attention_bwd_bf16 on the two CLIP sequence shapes (CUDA events, 20 launches after 3 warm-ups).
Default: the tcgen05 kernel (attention_bwd_tc.cu).  FC_ATTENTION_BWD=mma selects the mma.sync kernel of train.cu, whose
configuration FC_ATTN_BWD_CFG=10*warps+min_blocks overrides (both read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitclip_b200 import ops, train_ops as T  # noqa: E402

dev = torch.device("cuda:0")
for seqs, L, heads, causal in ((256, 197, 12, False), (2048, 197, 12, False), (512, 77, 8, True)):
    qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()
    dout = torch.randn(seqs * L, heads * 64, device=dev).bfloat16()
    out = ops.attention_bf16(qkv, seqs, L, heads, causal)
    dqkv = torch.empty_like(qkv)
    for _ in range(3):
        T.attention_bwd(qkv, out, dout, seqs, L, heads, causal, dqkv=dqkv)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        T.attention_bwd(qkv, out, dout, seqs, L, heads, causal, dqkv=dqkv)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    flops = 10.0 * L * L * 64 * heads * seqs
    which = os.environ.get("FC_ATTENTION_BWD", "tc") + ("/" + os.environ["FC_ATTN_BWD_CFG"] if "FC_ATTN_BWD_CFG" in os.environ else "")
    print(f"[{which}] attention_bwd seqs={seqs} L={L} heads={heads} "
          f"causal={causal}: {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (5 products)")
