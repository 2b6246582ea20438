"""tcgen05 GEMM parity (through the C ABI) against a plain PyTorch fp32 reference of the same op.

Tolerance: operands are the same bf16 bits on both sides and accumulation is fp32, so the only differences are
summation order (~1e-6 relative) and the final bf16 rounding of the output (2^-9 relative): atol 2e-2 + rtol 1e-2 for
bf16 outputs of O(1) magnitude, 1e-3 for fp32 outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, dev, seed=0, lda=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(M, lda or K, generator=g) * 0.5).to(dev).bfloat16()[:, :K]
    b = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    return a, b, bias


SHAPES = [
    (128, 256, 64),      # one tile, one k-block
    (128, 256, 256),     # pipeline wraps the 4 stages
    (256, 512, 768),     # several tiles
    (197, 768, 768),     # ragged M (TMA zero fill + row guard)
    (1000, 2304, 768),   # QKV-like, ragged M
    (77 * 3, 2048, 512),  # text fc1
    (50432, 768, 768),   # full vision pass height: 394 x 3 tiles over 148 persistent CTAs
    (4096, 768, 3072),   # fc2-like long K
    (300, 96, 128),      # N smaller than the tile
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bias(dev, M, N, K):
    from fitclip_b200 import ops
    a, b, bias = _mk(M, N, K, dev)
    out = ops.gemm_bf16(a, b, bias)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().T + bias
    err = (out.float() - ref).abs().max().item()
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=1e-2), f"max abs err {err}"


@pytest.mark.parametrize("M,N,K", [(197 * 4, 3072, 768), (77 * 5, 2048, 512)])
def test_gemm_quickgelu(dev, M, N, K):
    from fitclip_b200 import _lib, ops
    a, b, bias = _mk(M, N, K, dev, seed=1)
    out = ops.gemm_bf16(a, b, bias, epilogue=_lib.EPI_BIAS_QGELU)
    x = a.float() @ b.float().T + bias
    ref = x * torch.sigmoid(1.702 * x)
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=1e-2), (out.float() - ref).abs().max().item()


@pytest.mark.parametrize("M,N,K", [(197 * 4, 768, 768), (197 * 4, 768, 3072), (77 * 5, 512, 2048)])
def test_gemm_residual_in_place(dev, M, N, K):
    from fitclip_b200 import _lib, ops
    a, b, bias = _mk(M, N, K, dev, seed=2)
    x = torch.randn(M, N, device=dev).bfloat16()
    ref = x.float() + a.float() @ b.float().T + bias
    out = ops.gemm_bf16(a, b, bias, resid=x, epilogue=_lib.EPI_BIAS_RESID, out=x)  # C aliases the residual
    assert out.data_ptr() == x.data_ptr()
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=1e-2), (out.float() - ref).abs().max().item()


@pytest.mark.parametrize("M,N,K", [(1000, 1000, 512), (101, 37, 512), (333, 1000, 1536)])
def test_gemm_f32_out_ragged(dev, M, N, K):
    from fitclip_b200 import _lib, ops
    a, b, _ = _mk(M, N, K, dev, seed=3)
    out = ops.gemm_bf16(a, b, epilogue=_lib.EPI_F32, alpha=2.0)
    ref = 2.0 * (a.float() @ b.float().T)
    assert out.dtype == torch.float32
    assert torch.allclose(out, ref, atol=1e-3, rtol=1e-3), (out - ref).abs().max().item()


def test_gemm_strided_operand(dev):
    from fitclip_b200 import ops
    a, b, bias = _mk(512, 256, 64 * 3, dev, seed=4, lda=64 * 3 + 64)  # row stride > K
    out = ops.gemm_bf16(a, b, bias)
    ref = a.float() @ b.float().T + bias
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=1e-2)


def test_gemm_deterministic(dev):
    from fitclip_b200 import _lib, ops
    a, b, _ = _mk(2000, 1000, 512, dev, seed=5)
    o1 = ops.gemm_bf16(a, b, epilogue=_lib.EPI_F32)
    o2 = ops.gemm_bf16(a, b, epilogue=_lib.EPI_F32)
    assert torch.equal(o1, o2)


def test_gemm_rejects_bad_arguments(dev):
    from fitclip_b200 import _lib, ops
    a, b, bias = _mk(128, 256, 64, dev)
    with pytest.raises(_lib.FitclipError):
        ops.gemm_bf16(a, b, None)  # bias epilogue without a bias
    with pytest.raises(_lib.FitclipError):
        ops.gemm_bf16(a[:, :60], b[:, :60], bias)  # K not a multiple of 8
