"""End-to-end encoder parity: B200ClipVideoTextEncoder (bf16 tensor-core path) vs the fp32 CPU oracle, same seeded
weights and inputs.  Weights are the oracle's trained-like perturbation (no LayerNorm gamma / beta or attention bias at its identity
value), so the LayerNorm folding (W diag(gamma), b + W beta, column sums) and the bias adds run with real values.
Tolerances: cosine >= 0.999 per vector and, tighter, ~2-3x what was measured on the B200 /
what a bf16 residual stream explains (oracle.bf16_stream_model: the fp32 oracle with ONLY the stream rounded to bf16):
max-abs <= 5e-3 on unit-norm embeddings, mean-centred relative L2 error <= 0.1 (random-init embeddings are nearly
collinear, so the centred error is the sensitive number), and error(CUDA) <= 3 x error(bf16-stream emulation) + 1e-3."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _report(name, got, ref):
    cos = F.cosine_similarity(got, ref).min().item()
    max_abs = (got - ref).abs().max().item()
    gc, rc = got - got.mean(0, keepdim=True), ref - ref.mean(0, keepdim=True)
    centred = ((gc - rc).norm() / rc.norm()).item()
    print(f"{name}: min cos {cos:.6f}  max abs {max_abs:.3e}  centred rel L2 {centred:.3e}")
    return cos, max_abs, centred


@pytest.fixture(scope="module")
def models(dev):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    ref_model = oracle.clip_vit_b_16(seed=0)  # full ViT-B/16, random init
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(ref_model))
    enc = B200ClipVideoTextEncoder(ref_model.state_dict(), num_frames=4).to(dev)
    return ref, enc


def test_state_dict_layout_matches_openai_names(models):
    ref, enc = models
    assert [k for k, _ in enc.named_parameters()] == [k for k, _ in ref.named_parameters()]
    assert not any(k.endswith("logit_scale") for k in enc.state_dict())
    assert len(list(enc.parameters())) == 301


def test_encode_video_matches_oracle(models, dev):
    ref, enc = models
    g = torch.Generator().manual_seed(1234)
    video = torch.randn(6, 4, 3, 224, 224, generator=g)
    with torch.inference_mode():
        expect = ref.encode_video(video)
        got = enc.encode_video(video.to(dev)).cpu()
    assert got.shape == (6, 512) and got.dtype == torch.float32
    cos, max_abs, centred = _report("video", got, expect)
    assert cos >= 0.9995
    assert max_abs <= 5e-3
    assert centred <= 0.1


def test_encode_text_matches_oracle(models, dev):
    import oracle
    ref, enc = models
    ids = torch.cat([oracle.tokenize_synthetic(12, (4, 40), seed=4321), oracle.tokenize_synthetic(4, 77, seed=5)])
    with torch.inference_mode():
        expect = ref.encode_text({"input_ids": ids})
        got = enc.encode_text({"input_ids": ids.to(dev)}).cpu()
    cos, max_abs, centred = _report("text", got, expect)
    assert cos >= 0.9995
    assert max_abs <= 5e-3
    assert centred <= 0.1
    assert torch.allclose(got.norm(dim=-1), torch.ones(16), atol=1e-5)


def _truncate(state_dict, vision_layers, text_layers):
    """The first `vision_layers` / `text_layers` residual blocks of a CLIP state dict (heads and embeddings kept)."""
    out = {}
    for k, v in state_dict.items():
        if ".resblocks." in k:
            idx = int(k.split(".resblocks.")[1].split(".")[0])
            if idx >= (vision_layers if k.startswith("visual.") else text_layers):
                continue
        out[k] = v
    return out


@pytest.mark.parametrize("stress", [False, True], ids=["trained_like", "stress"])
def test_per_depth_activations_match_oracle(dev, stress):
    """SURVEY.md section 7 hard part 5 (per-layer activation checks): the residual stream after k blocks, read through
    ln_post / ln_final + projection, for k = 1, 2, 4, 8, 12 -- so an error that a later block would wash out (or that
    only accumulates with depth) is seen where it arises.  Weights are the trained-like perturbation (no LayerNorm or
    attention bias at its identity value); `stress` adds outlier channels, x10 gammas and +20 DC rows."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    full = oracle.clip_vit_b_16(seed=2, stress=stress).state_dict()
    video = torch.randn(3, 2, 3, 224, 224, generator=torch.Generator().manual_seed(21))
    ids = oracle.tokenize_synthetic(6, (4, 77), seed=22)
    worst = {}
    for depth in (1, 2, 4, 8, 12):
        sd = _truncate(full, depth, depth)
        model = oracle.build_model(sd)
        ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(model), num_frames=2)
        emu = oracle.RefClipVideoTextEncoder(oracle.bf16_stream_model(model), num_frames=2)
        enc = B200ClipVideoTextEncoder(sd, num_frames=2).to(dev)
        with torch.inference_mode():
            ev, et = ref(video, {"input_ids": ids})
            mv, mt = emu(video, {"input_ids": ids})
            gv, gt = enc(video.to(dev), {"input_ids": ids.to(dev)})
        tag = f"depth {depth}{' (stress)' if stress else ''}"
        cv, ct = _report(f"{tag} video", gv.cpu(), ev), _report(f"{tag} text", gt.cpu(), et)
        bv, bt = _report(f"{tag} video, bf16-stream emulation", mv, ev), _report(f"{tag} text, bf16-stream emulation", mt, et)
        worst[depth] = (cv, ct)
        assert cv[0] >= 0.9995 and ct[0] >= 0.9995, (depth, cv, ct)
        assert cv[1] <= 5e-3 and ct[1] <= 5e-3, (depth, cv, ct)
        assert cv[2] <= 0.1 and ct[2] <= 0.1, (depth, cv, ct)
        assert cv[1] <= 3 * bv[1] + 1e-3 and ct[1] <= 3 * bt[1] + 1e-3, (depth, cv, bv, ct, bt)


def test_stress_weights_full_depth(dev):
    """Pretrained-like stress case at full ViT-B/16 depth: residual outlier channels (x60 writers, x10 gammas), +20 DC
    offset on three token rows of each tower, random gamma / beta / attention biases everywhere."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    model = oracle.clip_vit_b_16(seed=5, stress=True)
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(model))
    emu = oracle.RefClipVideoTextEncoder(oracle.bf16_stream_model(model))
    enc = B200ClipVideoTextEncoder(model.state_dict(), num_frames=4).to(dev)
    video = torch.randn(4, 4, 3, 224, 224, generator=torch.Generator().manual_seed(31))
    ids = oracle.tokenize_synthetic(12, (4, 77), seed=32)
    with torch.inference_mode():
        ev, et = ref(video, {"input_ids": ids})
        mv, mt = emu(video, {"input_ids": ids})
        gv, gt = enc(video.to(dev), {"input_ids": ids.to(dev)})
    cv, ct = _report("stress video", gv.cpu(), ev), _report("stress text", gt.cpu(), et)
    bv, bt = _report("stress video, bf16-stream emulation", mv, ev), _report("stress text, bf16-stream emulation", mt, et)
    assert cv[0] >= 0.9995 and ct[0] >= 0.9995
    assert cv[1] <= 8e-3 and ct[1] <= 8e-3
    assert cv[2] <= 0.1 and ct[2] <= 0.1
    assert cv[1] <= 3 * bv[1] + 1e-3 and ct[1] <= 3 * bt[1] + 1e-3


def test_forward_tuple_and_int64_ids_and_bf16_frames(models, dev):
    import oracle
    ref, enc = models
    video = torch.randn(2, 4, 3, 224, 224, generator=torch.Generator().manual_seed(7)).to(dev)
    ids = oracle.tokenize_synthetic(2, (5, 20), seed=8).to(dev)
    v, t = enc(video, {"input_ids": ids})
    v2, t2 = enc(video=video.bfloat16(), text={"input_ids": ids.long()})
    assert torch.equal(t, t2)
    assert F.cosine_similarity(v, v2).min().item() > 0.9999


def test_passes_split_whole_videos(dev):
    """More videos than one internal pass holds: results must not depend on the pass boundary."""
    import oracle
    from fitclip_b200 import B200Clip, B200ClipVideoTextEncoder
    sd = oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1).state_dict()
    big = B200ClipVideoTextEncoder(B200Clip(sd, max_frames_per_pass=64, max_texts_per_pass=64)).to(dev)
    small = B200ClipVideoTextEncoder(B200Clip(sd, max_frames_per_pass=6, max_texts_per_pass=5)).to(dev)
    video = torch.randn(7, 3, 3, 224, 224, generator=torch.Generator().manual_seed(9)).to(dev)
    ids = oracle.tokenize_synthetic(13, (3, 77), seed=10).to(dev)
    assert torch.equal(big.encode_video(video), small.encode_video(video))
    assert torch.equal(big.encode_text({"input_ids": ids}), small.encode_text({"input_ids": ids}))
    # empty batches
    assert big.encode_video(video[:0]).shape == (0, 512)
    assert big.encode_text({"input_ids": ids[:0]}).shape == (0, 512)


def test_wise_matches_reference_and_reuploads_weights(dev):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, wise
    m1 = oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1)
    m2 = oracle.clip_vit_b_16(seed=1, vision_layers=1, transformer_layers=1)
    r1, r2 = oracle.RefClipVideoTextEncoder(copy.deepcopy(m1)), oracle.RefClipVideoTextEncoder(copy.deepcopy(m2))
    e1 = B200ClipVideoTextEncoder(m1.state_dict()).to(dev)
    e2 = B200ClipVideoTextEncoder(m2.state_dict()).to(dev)
    for w in (0.4, 0.5):  # config/encoder/wise.yaml:9 and aligner/wise.py:10
        ref = oracle.ref_wise(r1, r2, weight_for_2=w)
        got = wise(e1, e2, weight_for_2=w)
        assert type(got) is type(e1)
        ref_sd, got_sd = ref.state_dict(), got.state_dict()
        assert list(ref_sd) == list(got_sd)
        for k in ref_sd:
            assert torch.equal(ref_sd[k], got_sd[k].cpu()), k  # bit-exact lerp
        ids = oracle.tokenize_synthetic(3, (5, 30), seed=11)
        with torch.inference_mode():
            expect = ref.encode_text({"input_ids": ids})
        out = got.encode_text({"input_ids": ids.to(dev)}).cpu()
        assert F.cosine_similarity(out, expect).min().item() >= 0.999
        # and it differs from model1's output: the native weight copy really was refreshed
        assert not torch.allclose(out, e1.encode_text({"input_ids": ids.to(dev)}).cpu(), atol=1e-4)


def test_bad_inputs_fail_loudly(models, dev):
    from fitclip_b200 import _lib
    ref, enc = models
    with pytest.raises(_lib.FitclipError):
        enc.encode_video(torch.zeros(1, 1, 3, 224, 224))  # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        enc.encode_video(torch.zeros(1, 1, 3, 32, 32, device=dev))
    bad = torch.full((1, 77), 60000, dtype=torch.int32, device=dev)
    enc.encode_text({"input_ids": bad})
    with pytest.raises(_lib.FitclipError):
        enc.model.check_inputs()


GEOMETRIES = {
    # config/encoder/clip_vit_b_32.yaml: 224 / 32 -> 49 + 1 = 50 image tokens, 3072-wide patches
    "vit_b_32": dict(vision_patch_size=32, vision_layers=2, transformer_layers=2),
    # a narrow model: every width / head count / sequence length differs from ViT-B/16 (101 image tokens, 32 text tokens)
    "narrow": dict(embed_dim=256, image_resolution=160, vision_width=512, vision_layers=2, context_length=32,
                   transformer_width=256, transformer_heads=4, transformer_layers=2, vocab_size=1000),
    # config/encoder/clip_vit_l_14.yaml: 14-pixel patches (588-wide patch rows, padded to 592), 257 image tokens,
    # width 1024 / 16 heads, 768-wide text tower, 768-d embeddings
    "vit_l_14": dict(embed_dim=768, vision_patch_size=14, vision_width=1024, vision_layers=2, transformer_width=768,
                     transformer_heads=12, transformer_layers=1),
    # config/encoder/clip_vit_l_14_336px.yaml: 336 / 14 -> 576 + 1 = 577 image tokens
    "vit_l_14_336": dict(embed_dim=768, image_resolution=336, vision_patch_size=14, vision_width=1024, vision_layers=1,
                         transformer_width=768, transformer_heads=12, transformer_layers=1),
}


@pytest.mark.parametrize("name", sorted(GEOMETRIES))
def test_other_clip_geometries(dev, name):
    """the kernels are shape-generic (widths multiples of 64 up to 1024, head dim 64, sequences up
    to 768 tokens, any patch size dividing the resolution); the geometry is inferred from the state dict like
    clip.build_model."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    cfg = GEOMETRIES[name]
    ref_model = oracle.clip_vit_b_16(seed=3, **cfg)
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(ref_model))
    enc = B200ClipVideoTextEncoder(ref_model.state_dict(), num_frames=2).to(dev)
    res = cfg.get("image_resolution", 224)
    ctx = cfg.get("context_length", 77)
    vocab = cfg.get("vocab_size", 49408)
    video = torch.randn(5, 2, 3, res, res, generator=torch.Generator().manual_seed(11))
    ids = oracle.tokenize_synthetic(9, (3, ctx), seed=12, context_length=ctx, vocab_size=vocab)
    with torch.inference_mode():
        ev, et = ref(video, {"input_ids": ids})
        gv, gt = enc(video.to(dev), {"input_ids": ids.to(dev)})
    cos_v, max_v, _ = _report(f"{name} video", gv.cpu(), ev)
    cos_t, max_t, _ = _report(f"{name} text", gt.cpu(), et)
    assert cos_v >= 0.9995 and cos_t >= 0.9995
    assert max_v <= 5e-3 and max_t <= 5e-3


def test_unsupported_geometry_is_refused(dev):
    """8-pixel patches at 224 px give 785 image tokens, beyond the 768 the attention kernels take: the error must say so
    instead of computing something else."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, _lib
    model = oracle.clip_vit_b_16(seed=0, vision_patch_size=8, vision_layers=1, transformer_layers=1)
    enc = B200ClipVideoTextEncoder(model.state_dict()).to(dev)
    with pytest.raises(_lib.FitclipError, match="sequence length"):
        enc.encode_video(torch.zeros(1, 1, 3, 224, 224, device=dev))


def test_parameters_created_under_inference_mode(dev):
    """wise() / load_state_dict may run inside torch.inference_mode() (the reference wraps evaluation in it,
    aligner/__main__.py:64-66): such parameters have no version counter and must still be picked up."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, wise
    cfg = dict(vision_layers=1, transformer_layers=1)
    with torch.inference_mode():
        a = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **cfg).state_dict()).to(dev)
        b = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1, **cfg).state_dict()).to(dev)
        merged = wise(a, b, weight_for_2=0.5)
        video = torch.randn(2, 2, 3, 224, 224, device=dev)
        out = merged.encode_video(video)
        assert out.shape == (2, 512) and torch.isfinite(out).all()
        assert not torch.equal(out, a.encode_video(video))


def test_vit_l_14_full_depth(dev):
    """Row f4 at full depth: ViT-L/14 (config/encoder/clip_vit_l_14.yaml: 24 x 1024-wide vision blocks, 257 image tokens
    on the tcgen05 key-block attention kernel, 588-wide patch rows padded to 592; 12 x 768-wide text blocks), trained-like
    weights, against the fp32 oracle and the bf16-stream emulation."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    cfg = dict(embed_dim=768, vision_patch_size=14, vision_width=1024, vision_layers=24, transformer_width=768,
               transformer_heads=12)
    model = oracle.clip_vit_b_16(seed=4, **cfg)
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(model), num_frames=2)
    emu = oracle.RefClipVideoTextEncoder(oracle.bf16_stream_model(model), num_frames=2)
    enc = B200ClipVideoTextEncoder(model.state_dict(), num_frames=2).to(dev)
    video = torch.randn(3, 2, 3, 224, 224, generator=torch.Generator().manual_seed(41))
    ids = oracle.tokenize_synthetic(6, (4, 77), seed=42)
    with torch.inference_mode():
        ev, et = ref(video, {"input_ids": ids})
        mv, mt = emu(video, {"input_ids": ids})
        gv, gt = enc(video.to(dev), {"input_ids": ids.to(dev)})
    cv, ct = _report("ViT-L/14 video", gv.cpu(), ev), _report("ViT-L/14 text", gt.cpu(), et)
    bv, bt = _report("ViT-L/14 video, bf16-stream emulation", mv, ev), _report("ViT-L/14 text, bf16-stream emulation", mt, et)
    assert cv[0] >= 0.9995 and ct[0] >= 0.9995
    assert cv[1] <= 5e-3 and ct[1] <= 5e-3
    assert cv[1] <= 3 * bv[1] + 1e-3 and ct[1] <= 3 * bt[1] + 1e-3
