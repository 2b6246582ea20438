"""Parity of the memory-bound kernels and the fused attention against PyTorch fp32 references (via the C ABI)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D", [(1, 768), (197 * 3 + 5, 768), (77 * 7, 512), (33, 1024), (10, 64)])
def test_layernorm(dev, rows, D):
    from fitclip_b200 import ops
    torch.manual_seed(0)
    x = (torch.randn(rows, D, device=dev) * 3 + 1).bfloat16()
    g = torch.randn(D, device=dev)
    b = torch.randn(D, device=dev)
    out = ops.layernorm_bf16(x, g, b)
    ref = F.layer_norm(x.float(), (D,), g, b, 1e-5)
    # output rounding to bf16 dominates: 2^-9 relative
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=8e-3), (out.float() - ref).abs().max().item()
    # in place (ln_pre)
    y = x.clone()
    ops.layernorm_bf16(y, g, b, out=y)
    assert torch.equal(y, out)


def _ref_attention(qkv, seqs, L, heads, causal):
    D = heads * 64
    q, k, v = qkv.float().view(seqs, L, 3, heads, 64).permute(2, 0, 3, 1, 4)  # (3, S, H, L, 64)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=qkv.device).triu_(1)
    o = torch.softmax(s, dim=-1) @ v
    return o.permute(0, 2, 1, 3).reshape(seqs * L, D)


@pytest.mark.parametrize("seqs,L,heads,causal", [
    (3, 197, 12, False), (5, 77, 8, True), (2, 77, 8, False), (4, 16, 2, True), (3, 50, 4, False),
    (2, 130, 3, True), (1, 208, 1, False), (7, 1, 2, True),
    # many items per persistent CTA: exercises the S(i+1) / O(i) overlap of the tcgen05 kernel
    (70, 197, 12, False), (33, 193, 12, False), (40, 208, 6, False), (61, 200, 5, False),
    # long sequences (key blocks of 64 streamed past a 128-row query chunk): ViT-L/14 257, ViT-L/14@336 577
    (3, 257, 16, False), (2, 577, 4, False), (2, 300, 2, True), (1, 768, 1, True), (2, 209, 3, False),
    # the causal text sequence on the tcgen05 kernel (65..80 tokens), many items per persistent CTA
    (300, 77, 8, True), (9, 65, 2, True), (11, 80, 3, True), (64, 64, 2, True),
    # every length up to 208 runs on the tcgen05 kernels (padded key counts 80 / 144 / 208), masked or not:
    # ViT-B/32's 50-token image sequence, lengths at and around the padding boundaries
    (200, 50, 12, False), (6, 33, 2, False), (9, 100, 3, False), (9, 144, 2, False), (5, 145, 2, False), (4, 192, 3, False),
    (5, 145, 2, True), (3, 160, 2, True), (2, 208, 2, True), (6, 81, 1, True), (5, 48, 2, True),
    # tcgen05 key-block kernel (un-masked 209..768 tokens): many items per CTA, exact multiples of the block, the maximum
    (40, 257, 16, False), (10, 577, 16, False), (3, 768, 2, False), (5, 256, 4, False), (6, 385, 3, False), (4, 384, 2, False)])
def test_attention(dev, seqs, L, heads, causal):
    from fitclip_b200 import ops
    torch.manual_seed(1)
    qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()
    out = ops.attention_bf16(qkv, seqs, L, heads, causal)
    ref = _ref_attention(qkv, seqs, L, heads, causal)
    # P is rounded to bf16 before PV and the output is bf16: ~1e-2 absolute on O(1) values
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=2e-2), (out.float() - ref).abs().max().item()


@pytest.mark.parametrize("B,T,D", [(1, 1, 512), (32, 4, 512), (7, 8, 512), (1000, 1, 512), (3, 5, 100)])
def test_pool_normalize(dev, B, T, D):
    from fitclip_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(B * T, D, device=dev) * 5
    out = ops.pool_normalize(x, T)
    ref = (x / x.norm(dim=-1, keepdim=True)).view(B, T, D).mean(dim=1)  # clip_video_text_encoder.py:85-89
    assert (out - ref).abs().max().item() <= 1e-6


@pytest.mark.parametrize("n", [1, 3, 4, 1000, 768 * 3072 + 3, 49408 * 512])
@pytest.mark.parametrize("w", [0.4, 0.5, 0.0, 1.0, 0.123456789])
def test_wise_lerp_bit_exact(dev, n, w):
    from fitclip_b200 import ops
    torch.manual_seed(3)
    p1 = torch.randn(n, device=dev)
    p2 = torch.randn(n, device=dev)
    out = ops.wise_lerp(p1, p2, w)
    ref = (1 - w) * p1.cpu() + w * p2.cpu()  # aligner/wise.py:16 evaluated by torch on the CPU
    assert torch.equal(out.cpu(), ref)


def test_wise_lerp_bf16_copy(dev):
    from fitclip_b200 import ops
    p1 = torch.randn(1001, device=dev)
    p2 = torch.randn(1001, device=dev)
    ob = torch.empty(1001, device=dev, dtype=torch.bfloat16)
    out = ops.wise_lerp(p1, p2, 0.4, out_bf16=ob)
    assert torch.equal(ob, out.bfloat16())
