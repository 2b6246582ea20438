"""Parity of the training-step kernels (fitclip_b200/csrc/train.cu) against torch.autograd on the
same inputs, through the C ABI.  Tolerances: outputs are bf16 (2^-9 relative rounding) with fp32 accumulation; fp32
outputs (weight gradients, loss gradients, AdamW) are compared at fp32-level tolerances."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(a, b, rel, what="", floor=1e-6):
    """max abs error within `rel` of the reference's largest magnitude (+ an absolute floor for all-zero references,
    e.g. dq / dk of a one-token sequence)."""
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= rel * ref + floor, f"{what}: max abs err {err:.3e} vs scale {ref:.3e}"


@pytest.mark.parametrize("M,N,K,splits", [(768, 2304, 197 * 9 + 3, 0), (256, 256, 64, 1), (130, 100, 1000, 7),
                                          (3072, 768, 8192, 0), (512, 512, 40, 0)])
def test_gemm_splitk(dev, M, N, K, splits):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(0)
    ld = T.pad8(K)
    a = torch.zeros(M, ld, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(N, ld, device=dev, dtype=torch.bfloat16)
    a[:, :K] = torch.randn(M, K, device=dev)
    b[:, :K] = torch.randn(N, K, device=dev)
    out = torch.zeros(M, N, device=dev)
    T.gemm_splitk(a, b, out, alpha=0.5, k=ld, k_splits=splits)
    ref = 0.5 * (a.float() @ b.float().T)
    _close(out, ref, 2e-5, "split-K GEMM")
    T.gemm_splitk(a, b, out, alpha=0.5, k=ld, k_splits=splits)  # accumulates
    _close(out, 2 * ref, 2e-5, "split-K GEMM (accumulate)")


@pytest.mark.parametrize("rows,cols,gl,gs", [(197 * 5, 768, 0, 0), (197 * 5, 768, 197, 1), (77, 64, 0, 0),
                                             (1000, 3072, 0, 0), (50 * 3, 128, 50, 1), (7, 8, 0, 0)])
def test_transpose_colsum(dev, rows, cols, gl, gs):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(1)
    x = torch.randn(rows, cols, device=dev).bfloat16()
    cs = torch.ones(cols, device=dev)
    out = T.transpose(x, gl, gs, colsum=cs)
    kept = x if gl == 0 else x.view(rows // gl, gl, cols)[:, gs:].reshape(-1, cols)
    assert out.shape == (cols, T.pad8(kept.shape[0]))
    assert torch.equal(out[:, :kept.shape[0]], kept.T)
    assert (out[:, kept.shape[0]:] == 0).all()
    _close(cs, 1 + kept.float().sum(0), 1e-5, "colsum")


@pytest.mark.parametrize("rows,D,with_add", [(197 * 3 + 5, 768, True), (77 * 7, 512, False), (33, 1024, True),
                                             (10, 64, False), (5000, 768, True)])
def test_layernorm_bwd(dev, rows, D, with_add):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(2)
    x = (torch.randn(rows, D, device=dev) * 2 + 0.5).bfloat16()
    dy = torch.randn(rows, D, device=dev).bfloat16()
    g = torch.randn(D, device=dev)
    b = torch.randn(D, device=dev)
    add = torch.randn(rows, D, device=dev).bfloat16() if with_add else None
    xf = x.float().requires_grad_(True)
    gf, bf = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xf, (D,), gf, bf, 1e-5).backward(dy.float())
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    dx = T.layernorm_bwd(x, dy, g, dg, db, add=add)
    ref = xf.grad + (add.float() if with_add else 0)
    _close(dx, ref, 8e-3, "dx")
    _close(dg, gf.grad, 1e-4, "dgamma")
    _close(db, bf.grad, 1e-4, "dbeta")
    if with_add:  # in place over `add`
        T.layernorm_bwd(x, dy, g, dg, db, add=add, out=add)
        assert torch.equal(add, dx)


def test_quickgelu(dev):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(3)
    u = (torch.randn(1000, 3072, device=dev) * 2).bfloat16()
    dg = torch.randn(1000, 3072, device=dev).bfloat16()
    uf = u.float().requires_grad_(True)
    ref = uf * torch.sigmoid(1.702 * uf)
    ref.backward(dg.float())
    _close(T.quickgelu(u), ref.detach(), 5e-3, "quickgelu")
    du = T.quickgelu_bwd(u, dg)
    _close(du, uf.grad, 5e-3, "quickgelu backward")
    g = torch.empty_like(u)
    T.quickgelu_bwd(u, dg, out=dg, g_out=g)
    assert torch.equal(dg, du)
    _close(g, ref.detach(), 5e-3, "quickgelu from the backward pass")


def _ref_attention(qkv, seqs, L, heads, causal):
    D = heads * 64
    q, k, v = qkv.view(seqs, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=qkv.device).triu_(1)
    o = torch.softmax(s, dim=-1) @ v
    return o.permute(0, 2, 1, 3).reshape(seqs * L, D)


@pytest.mark.parametrize("seqs,L,heads,causal", [
    (3, 197, 12, False), (5, 77, 8, True), (2, 77, 8, False), (4, 16, 2, True), (3, 50, 4, False), (2, 130, 3, True),
    (1, 208, 1, False), (7, 1, 2, True), (2, 257, 2, False), (3, 33, 1, True), (40, 197, 3, False),
    # tile / column-layout edges of the tcgen05 kernel: exactly one tile, one row in the second tile, NP = 64 (the
    # accumulators start at NP + 64 below that), causal across two tiles, many items
    (2, 128, 2, False), (2, 129, 2, True), (3, 64, 1, True), (2, 17, 3, True), (2, 200, 2, True), (3, 48, 2, False),
    (300, 77, 8, True)])
def test_attention_bwd(dev, seqs, L, heads, causal):
    from fitclip_b200 import ops, train_ops as T
    torch.manual_seed(4)
    qkv = torch.randn(seqs * L, 3 * heads * 64, device=dev).bfloat16()
    dout = torch.randn(seqs * L, heads * 64, device=dev).bfloat16()
    out = ops.attention_bf16(qkv, seqs, L, heads, causal)
    qf = qkv.float().requires_grad_(True)
    ref_out = _ref_attention(qf, seqs, L, heads, causal)
    ref_out.backward(dout.float())
    dqkv = T.attention_bwd(qkv, out, dout, seqs, L, heads, causal)
    D = heads * 64
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        _close(dqkv[:, sl], qf.grad[:, sl], 2e-2, name)


@pytest.mark.parametrize("B", [1, 7, 64, 512])
def test_loss_fwd_bwd(dev, B):
    from fitclip_b200 import train_ops as T
    from oracle.loss_ref import ref_nce_loss, ref_teacher_student_nce_loss
    torch.manual_seed(5)
    s = (torch.randn(B, B, device=dev) * 3).requires_grad_(True)
    t = torch.randn(B, B, device=dev) * 3
    ref = ref_nce_loss(s)
    ref.backward()
    loss, grad = T.loss_fwd_bwd(s.detach(), None, gscale=0.5)
    assert torch.allclose(loss, ref.detach(), rtol=1e-5, atol=1e-5)
    _close(grad, 0.5 * s.grad, 2e-5, "nce grad")
    s.grad = None
    ref = ref_teacher_student_nce_loss(s, t, reduction="batchmean")
    ref.backward()
    loss, grad = T.loss_fwd_bwd(s.detach(), t, gscale=2.0)
    assert torch.allclose(loss, ref.detach(), rtol=2e-5, atol=2e-5)
    _close(grad, 2.0 * s.grad, 2e-5, "ts grad")


@pytest.mark.parametrize("M,N,K", [(197 * 9, 3072, 768), (300, 256, 128), (77 * 5 + 3, 2048, 512)])
def test_dgrad_fused_with_quickgelu_backward(dev, M, N, K):
    """``(dy @ W) * quickgelu'(u)`` from the dgrad GEMM's epilogue (EPI_QGELU_BWD, B read MN-major) against autograd of
    ``quickgelu(u) @ W.T`` w.r.t. u, and against the two-kernel path it replaces."""
    from fitclip_b200 import train_ops as T
    torch.manual_seed(M + N)
    dy = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(K, N, device=dev) * K ** -0.5).bfloat16()   # c_proj.weight as stored: (out = K, in = N)
    u = (torch.randn(M, N, device=dev) * 2).bfloat16()
    uf = u.float().requires_grad_(True)
    (uf * torch.sigmoid(1.702 * uf) @ w.float().T * dy.float()).sum().backward()
    got = T.gemm_nt(dy, w, None, qgelu_bwd_of=u)
    _close(got, uf.grad, 2e-2, "fused du")
    two_step = T.quickgelu_bwd(u, T.gemm_nt(dy, w, torch.zeros(N, device=dev)))
    assert (got.float() - two_step.float()).abs().max().item() <= 2e-2 * uf.grad.abs().max().item()


@pytest.mark.parametrize("R,C", [(6, 5), (1, 9), (300, 17), (33, 600)])
def test_teacher_student_loss_on_rectangular_scores(dev, R, C):
    """(videos x prompts) matrices (teacher_student.py:104-120): ``batchmean`` divides the row direction by R and the
    column direction -- the loss of the transposed matrices, loss.py:36-39 -- by C."""
    from fitclip_b200 import train_ops as T
    from oracle.loss_ref import ref_teacher_student_nce_loss
    torch.manual_seed(R * 1000 + C)
    s = (torch.randn(R, C, device=dev) * 3).requires_grad_(True)
    t = torch.randn(R, C, device=dev) * 3
    ref = ref_teacher_student_nce_loss(s, t, reduction="batchmean")
    ref.backward()
    loss, grad = T.loss_fwd_bwd(s.detach(), t, gscale=1.5)
    assert torch.allclose(loss, ref.detach(), rtol=2e-5, atol=2e-5)
    _close(grad, 1.5 * s.grad, 2e-5, "ts grad (rectangular)")
    with pytest.raises(AssertionError):
        T.loss_fwd_bwd(s.detach(), None)  # nce_loss pairs row i with column i


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_sgemm(dev, ta, tb):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(6)
    M, N, K = 130, 77, 513
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev)
    ref = 0.3 * ((a.T if ta else a).double() @ (b.T if tb else b).double())
    _close(T.sgemm(a, b, ta, tb, alpha=0.3), ref, 1e-5, "sgemm")


@pytest.mark.parametrize("B,T_,D", [(5, 4, 512), (3, 1, 512), (2, 8, 768)])
def test_pool_normalize_bwd(dev, B, T_, D):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(7)
    x = torch.randn(B * T_, D, device=dev, requires_grad=True)
    dout = torch.randn(B, D, device=dev)
    out = (x / x.norm(dim=-1, keepdim=True)).view(B, T_, D).mean(1)
    out.backward(dout)
    dx = T.pool_normalize_bwd(x.detach(), dout, T_)
    _close(dx, x.grad, 8e-3, "pool+normalise backward")


def test_seq_rows_and_embedding_grads(dev):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(8)
    S, L, W, V = 9, 13, 128, 50
    x = torch.randn(S * L, W, device=dev).bfloat16()
    ids = torch.randint(1, V - 1, (S, L), device=dev, dtype=torch.int32)
    eot = torch.randint(1, L, (S,), device=dev)
    for s in range(S):
        ids[s, eot[s]] = V - 1
        ids[s, eot[s] + 1:] = 0
    rows = T.gather_seq_rows(x, ids, S, L)
    assert torch.equal(rows, x.view(S, L, W)[torch.arange(S), eot])
    assert torch.equal(T.gather_seq_rows(x, None, S, L), x.view(S, L, W)[:, 0])
    back = T.scatter_seq_rows(rows, ids, L).view(S, L, W)
    assert torch.equal(back[torch.arange(S), eot], rows) and back.float().abs().sum() == rows.float().abs().sum()
    pos = torch.zeros(L, W, device=dev)
    T.seq_sum(x, pos, S, L)
    _close(pos, x.float().view(S, L, W).sum(0), 1e-5, "seq_sum")
    dtok = torch.zeros(V, W, device=dev)
    T.token_scatter_add(ids.view(-1), x, dtok)
    ref = torch.zeros(V, W, device=dev).index_add_(0, ids.view(-1).long(), x.float())
    _close(dtok, ref, 1e-5, "token scatter")


def test_adamw(dev):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(9)
    n = 100003
    p = torch.randn(n, device=dev)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=3e-3)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    pb = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in range(1, 4):
        g = torch.randn(n, device=dev)
        ref_p.grad = g.clone()
        opt.step()
        T.adamw_step(p, g, m, v, step, lr=3e-3, p_bf16=pb)
        assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-6), (p - ref_p.detach()).abs().max().item()
    assert torch.equal(pb, p.bfloat16())


def test_patch_and_text_embed(dev):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(10)
    Fr, R, P, W = 3, 64, 16, 128
    frames = torch.randn(Fr, 3, R, R, device=dev)
    conv = torch.randn(W, 3, P, P, device=dev) * 0.05
    cls, pos = torch.randn(W, device=dev), torch.randn((R // P) ** 2 + 1, W, device=dev)
    x, patches = T.patch_embed(frames, conv.view(W, -1).bfloat16().contiguous(), cls, pos, P)
    ref = F.conv2d(frames.bfloat16().float(), conv.bfloat16().float(), stride=P).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.expand(Fr, 1, W), ref], 1) + pos
    _close(x.view(Fr, -1, W), ref, 8e-3, "patch embed")
    assert patches.shape == (Fr * (R // P) ** 2, 3 * P * P)
    V, L, C = 40, 9, 5
    tok, tpos = torch.randn(V, W, device=dev), torch.randn(L, W, device=dev)
    ids = torch.randint(0, V, (C, L), device=dev, dtype=torch.int32)
    err = torch.zeros(64, device=dev, dtype=torch.int32)
    xt = T.text_embed(ids, tok, tpos, err)
    _close(xt.view(C, L, W), tok[ids.long()] + tpos, 8e-3, "text embed")
    assert int(err[0]) == 0


@pytest.mark.parametrize("M,K,N,resid", [(197 * 3, 768, 3072, False), (544, 3072, 768, True), (100, 64, 64, False),
                                         (1000, 2304, 768, True), (33, 512, 768, False)])
def test_gemm_nt_reads_b_in_place(dev, M, K, N, resid):
    """dgrad dX = dY @ W with W (K, N) row-major read through an MN-major tcgen05 descriptor."""
    from fitclip_b200 import train_ops as T
    torch.manual_seed(11)
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(K, N, device=dev) / K ** 0.5).bfloat16()
    r = torch.randn(M, N, device=dev).bfloat16() if resid else None
    bias = torch.randn(N, device=dev)
    ref = a.float() @ w.float() + bias + (r.float() if resid else 0)
    _close(T.gemm_nt(a, w, bias, r), ref, 8e-3, "gemm_nt bf16")
    _close(T.gemm_nt(a, w, None, f32=True), a.float() @ w.float(), 1e-5, "gemm_nt fp32")


@pytest.mark.parametrize("rows,N,K", [(197 * 9 + 3, 2304, 768), (77, 64, 64), (1000, 768, 3072), (544, 768, 768),
                                      (4096 + 8, 512, 2048), (31, 512, 768)])
def test_wgrad_tn_reads_both_in_place(dev, rows, N, K):
    """wgrad dW = dY^T @ X with dY (rows, N) and X (rows, K) row-major, both MN-major operands, split-K."""
    from fitclip_b200 import train_ops as T
    torch.manual_seed(12)
    dy = torch.randn(rows, N, device=dev).bfloat16()
    x = torch.randn(rows, K, device=dev).bfloat16()
    out = torch.zeros(N, K, device=dev)
    T.wgrad_tn(dy, x, out, alpha=0.5)
    ref = 0.5 * (dy.float().T @ x.float())
    _close(out, ref, 2e-5, "wgrad_tn")
    T.wgrad_tn(dy, x, out, alpha=0.5, k_splits=3)
    _close(out, 2 * ref, 2e-5, "wgrad_tn accumulate")


@pytest.mark.parametrize("rows,cols", [(197 * 9 + 3, 2304), (1, 64), (100000, 768), (77, 136), (33, 8)])
def test_colsum(dev, rows, cols):
    from fitclip_b200 import train_ops as T
    torch.manual_seed(13)
    x = torch.randn(rows, cols, device=dev).bfloat16()
    out = torch.ones(cols, device=dev)
    T.colsum(x, out)
    _close(out, 1 + x.float().sum(0), 2e-5, "colsum", floor=1e-3)
