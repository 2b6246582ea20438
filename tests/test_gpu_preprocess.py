"""GPU eval pre-processing kernel (fc_preprocess_frames) against the CPU oracle of the reference transform
(oracle/preprocess_ref.py <- aligner/encoder/clip_video_text_encoder.py:124-133).

Tolerance: fp32 taps and weights on both sides, different summation order and FMA contraction: max-abs <= 2e-5 on
normalised pixels (|x| <= ~2.7, 1/std ~ 3.7 amplifies the interpolation rounding); bf16 output: one bf16 rounding
(2^-9 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


@pytest.mark.parametrize("h,w", [(240, 320), (360, 202), (224, 224), (256, 256), (113, 400), (480, 270), (1080, 1920)])
def test_preprocess_matches_oracle(dev, h, w):
    import oracle
    from fitclip_b200 import ops
    g = torch.Generator().manual_seed(h * 1000 + w)
    video = torch.randint(0, 256, (2, h, w, 3), dtype=torch.uint8, generator=g)
    ref = oracle.ref_eval_transform(video, 224, MEAN, STD)
    got = ops.preprocess_frames(video.to(dev), 224, MEAN, STD)
    assert got.shape == (2, 3, 224, 224) and got.dtype == torch.float32
    err = (got.cpu() - ref).abs().max().item()
    assert err <= 2e-5, err
    got16 = ops.preprocess_frames(video.to(dev), 224, MEAN, STD, torch.bfloat16)
    assert got16.dtype == torch.bfloat16
    assert torch.allclose(got16.float().cpu(), ref, atol=1e-5, rtol=2 ** -8)


@pytest.mark.parametrize("h,w", [(240, 320), (360, 202), (224, 224), (113, 400), (1080, 1920), (100, 90)])
def test_bilinear_preprocess_matches_oracle(dev, h, w):
    """The SLIP wrapper's transform (slip_video_text_encoder.py:78-87): bilinear resize (up- and down-scaling, borders),
    ImageNet statistics; NCHW output and the patch-matrix output of the fused path."""
    import oracle
    from fitclip_b200 import ops
    from fitclip_b200.slip_encoder import IMAGENET_MEAN, IMAGENET_STD
    g = torch.Generator().manual_seed(h * 1000 + w + 1)
    video = torch.randint(0, 256, (2, h, w, 3), dtype=torch.uint8, generator=g)
    ref = oracle.ref_eval_transform(video, 224, IMAGENET_MEAN, IMAGENET_STD, interpolation="bilinear")
    got = ops.preprocess_frames(video.to(dev), 224, IMAGENET_MEAN, IMAGENET_STD, interpolation="bilinear")
    err = (got.cpu() - ref).abs().max().item()
    assert err <= 2e-5, err
    patches = ops.preprocess_to_patches(video.to(dev), 224, 16, IMAGENET_MEAN, IMAGENET_STD, interpolation="bilinear")
    unfold = torch.nn.functional.unfold(ref, kernel_size=16, stride=16).transpose(1, 2).reshape(-1, 768)
    assert torch.allclose(patches.float().cpu(), unfold, atol=1e-5, rtol=2 ** -8)
    with pytest.raises(KeyError):
        ops.preprocess_frames(video.to(dev), 224, IMAGENET_MEAN, IMAGENET_STD, interpolation="nearest")


def test_slip_encode_video_uint8_equals_transform_then_encode(dev):
    import oracle
    from fitclip_b200 import B200SlipVideoTextEncoder
    enc = B200SlipVideoTextEncoder(oracle.slip_clip_vit_b_16(seed=2, vision_layers=1, transformer_layers=1).state_dict(),
                                   num_frames=2).to(dev)
    raw = torch.randint(0, 256, (3, 2, 180, 250, 3), dtype=torch.uint8, device=dev,
                        generator=torch.Generator(device=dev).manual_seed(4))
    with torch.inference_mode():
        fused = enc.encode_video_uint8(raw)
        two_step = enc.encode_video_uint8(raw, dtype=torch.bfloat16)
        hook = enc.get_eval_transform(torch.float32)
        host = enc.encode_video(torch.stack([hook(v.cpu()) for v in raw]).to(dev))
    assert torch.equal(fused, two_step)  # same bf16 pixels reach the patch GEMM either way
    assert torch.nn.functional.cosine_similarity(fused, host).min().item() >= 0.9999


def test_preprocess_batched_leading_dims_and_flat_colour(dev):
    from fitclip_b200 import ops
    # a constant image must stay constant through the cubic kernel (weights sum to one)
    video = torch.full((2, 3, 300, 260, 3), 128, dtype=torch.uint8, device=dev)
    out = ops.preprocess_frames(video, 224, MEAN, STD)
    assert out.shape == (2, 3, 3, 224, 224)
    for c in range(3):
        expect = (128 / 255 - MEAN[c]) / STD[c]
        assert (out[:, :, c] - expect).abs().max().item() <= 1e-5


def test_encode_video_uint8_equals_transform_then_encode(dev):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, ops
    sd = oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1).state_dict()
    enc = B200ClipVideoTextEncoder(sd, num_frames=2).to(dev)
    g = torch.Generator().manual_seed(5)
    raw = torch.randint(0, 256, (3, 2, 240, 320, 3), dtype=torch.uint8, generator=g).to(dev)
    a = enc.encode_video_uint8(raw, dtype=torch.float32)
    b = enc.encode_video(ops.preprocess_frames(raw, 224, MEAN, STD))
    assert torch.equal(a, b)
    # fused path (row f1 as specified): the transform writes bf16 patch rows straight into the patch-embedding GEMM's
    # operand -- bit-identical to transform -> bf16 NCHW frame -> im2col, for several frame sizes / patch sizes
    fused = enc.encode_video_uint8(raw)
    assert torch.equal(fused, enc.encode_video_uint8(raw, dtype=torch.bfloat16))
    tall = torch.randint(0, 256, (2, 2, 360, 250, 3), dtype=torch.uint8, generator=g).to(dev)
    assert torch.equal(enc.encode_video_uint8(tall), enc.encode_video_uint8(tall, dtype=torch.bfloat16))
    # and against the CPU path: hook on the CPU -> oracle encoder
    ref_enc = oracle.RefClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1))
    frames = torch.stack([oracle.ref_eval_transform(v, 224, MEAN, STD) for v in raw.cpu()])
    with torch.inference_mode():
        ref = ref_enc.encode_video(frames)
    cos = torch.nn.functional.cosine_similarity(a.cpu(), ref).min().item()
    assert cos >= 0.999, cos


def test_preprocess_rejects_bad_input(dev):
    from fitclip_b200 import _lib, ops
    with pytest.raises(ValueError):
        ops.preprocess_frames(torch.zeros(1, 8, 8, 3, device=dev), 224, MEAN, STD)  # not uint8
    with pytest.raises(_lib.FitclipError):
        ops.preprocess_frames(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), 224, MEAN, STD)  # CPU tensor


@pytest.mark.parametrize("patch", [16, 32, 14])
def test_preprocess_to_patches_equals_unfold_of_the_frames(dev, patch):
    """The fused transform + patch gather against torch's unfold of the two-step result (bit-exact: same arithmetic,
    another store address)."""
    from fitclip_b200 import ops
    g = torch.Generator().manual_seed(8)
    raw = torch.randint(0, 256, (5, 240, 320, 3), dtype=torch.uint8, generator=g).to(dev)
    frames = ops.preprocess_frames(raw, 224, MEAN, STD, torch.bfloat16)              # (5, 3, 224, 224)
    expect = torch.nn.functional.unfold(frames.float(), kernel_size=patch, stride=patch)  # (5, 3*P*P, G*G)
    expect = expect.transpose(1, 2).reshape(-1, 3 * patch * patch).bfloat16()
    got = ops.preprocess_to_patches(raw, 224, patch, MEAN, STD)
    assert got.shape == expect.shape
    bad = (got != expect).nonzero()
    assert bad.numel() == 0, (bad[:8].tolist(), got[bad[0, 0], bad[0, 1]].item(), expect[bad[0, 0], bad[0, 1]].item())


def test_fused_uint8_path_other_patch_sizes(dev):
    """ViT-B/32 (32-pixel patches) and ViT-L/14 (14-pixel patches, patch rows padded 588 -> 592 with zero columns)."""
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    g = torch.Generator().manual_seed(6)
    raw = torch.randint(0, 256, (2, 2, 256, 300, 3), dtype=torch.uint8, generator=g).to(dev)
    for cfg in (dict(vision_patch_size=32, vision_layers=1, transformer_layers=1),
                dict(embed_dim=768, vision_patch_size=14, vision_width=1024, vision_layers=1, transformer_width=768,
                     transformer_heads=12, transformer_layers=1)):
        enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=3, **cfg).state_dict(), num_frames=2).to(dev)
        assert torch.equal(enc.encode_video_uint8(raw), enc.encode_video_uint8(raw, dtype=torch.bfloat16)), cfg


def test_fused_path_argument_errors(dev):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, _lib, ops
    with pytest.raises(ValueError):
        ops.preprocess_to_patches(torch.zeros(1, 8, 8, 3, device=dev), 224, 16, MEAN, STD)          # not uint8
    with pytest.raises(_lib.FitclipError):
        ops.preprocess_to_patches(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), 224, 16, MEAN, STD)   # CPU tensor
    with pytest.raises(_lib.FitclipError, match="size % patch"):
        ops.preprocess_to_patches(torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device=dev), 224, 15, MEAN, STD)
    enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1).state_dict(),
                                   num_frames=2).to(dev)
    with pytest.raises(ValueError):
        enc.encode_video_uint8(torch.zeros(1, 2, 3, 32, 32, dtype=torch.uint8, device=dev))          # channels-first
    with pytest.raises(_lib.FitclipError):
        enc.encode_video_uint8(torch.zeros(1, 2, 32, 32, 3, dtype=torch.uint8))                       # CPU tensor
    assert enc.encode_video_uint8(torch.zeros(0, 2, 32, 32, 3, dtype=torch.uint8, device=dev)).shape == (0, 512)
