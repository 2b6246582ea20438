"""SLIP-layout models on the CUDA path (``config/encoder/slip*.yaml``, ``slip.py:595-600,618-623``):
timm-ViT vision tower (no ln_pre, exact GELU, LayerNorm eps 1e-6, patch-embedding bias) + CLIP text tower, against

* the outputs of the reference's own ``slip.CLIP`` / ``SlipVideoTextEncoder`` / ``wise`` (tests/golden/reference_slip.pt);
* the fp32 CPU oracle at full ViT-B/16 depth with trained-like LayerNorm / bias values, same bars as test_gpu_encoder.py
  (cosine >= 0.9995, max-abs <= 5e-3 on unit-norm embeddings, centred relative L2 <= 0.1);
* the GEMM epilogue with the exact GELU alone (fc_gemm_bf16 is not exposed for it, so through a one-block model)."""
import copy
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_slip.pt")


@pytest.fixture(scope="module")
def ref():
    return torch.load(GOLDEN, map_location="cpu", weights_only=False)


def _report(name, got, expect):
    cos = F.cosine_similarity(got, expect).min().item()
    max_abs = (got - expect).abs().max().item()
    gc, rc = got - got.mean(0, keepdim=True), expect - expect.mean(0, keepdim=True)
    centred = ((gc - rc).norm() / rc.norm()).item()
    print(f"{name}: min cos {cos:.6f}  max abs {max_abs:.3e}  centred rel L2 {centred:.3e}")
    return cos, max_abs, centred


def test_cuda_slip_encoder_matches_reference_wrapper(ref, dev):
    from fitclip_b200 import B200SlipVideoTextEncoder, load_slip_model
    enc = B200SlipVideoTextEncoder(load_slip_model({"state_dict": ref["checkpoint_1"]}), num_frames=3).to(dev)
    with torch.inference_mode():
        v, t = enc(ref["video"].to(dev), {"input_ids": ref["input_ids"].to(dev)})
        feats = enc.model.encode_image(ref["video"][:, 0].to(dev)).cpu()
    for got, expect in ((v.cpu(), ref["wrapper_video_emb"]), (t.cpu(), ref["wrapper_text_emb"])):
        assert F.cosine_similarity(got, expect).min().item() >= 0.999
        assert (got - expect).abs().max().item() <= 2e-2
    # slip.CLIP.encode_image, un-normalised (slip.py:462-466)
    assert F.cosine_similarity(feats, ref["image_features"]).min().item() >= 0.999
    assert ((feats - ref["image_features"]).norm(dim=1) / ref["image_features"].norm(dim=1)).max().item() <= 3e-2


def test_cuda_slip_wise_matches_reference(ref, dev):
    from fitclip_b200 import B200SlipVideoTextEncoder
    from fitclip_b200.wise import wise, wise_state_dict
    a = B200SlipVideoTextEncoder(ref["checkpoint_1"], num_frames=3).to(dev)
    b = B200SlipVideoTextEncoder(ref["checkpoint_2"], num_frames=3).to(dev)
    for w in (0.4, 0.5):
        ours = wise_state_dict(a, b, weight_for_2=w)
        for k, expect in ref[f"wise_{w}_state_dict"].items():
            assert torch.equal(ours[k].cpu(), expect), (w, k)  # bit-exact, under the checkpoint's own names
        merged = wise(a, b, weight_for_2=w)
        with torch.inference_mode():
            v, t = merged(ref["video"].to(dev), {"input_ids": ref["input_ids"].to(dev)})
        assert F.cosine_similarity(v.cpu(), ref[f"wise_{w}_video_emb"]).min().item() >= 0.999
        assert F.cosine_similarity(t.cpu(), ref[f"wise_{w}_text_emb"]).min().item() >= 0.999


@pytest.fixture(scope="module")
def full(dev):
    import oracle
    from fitclip_b200 import B200SlipVideoTextEncoder
    model = oracle.slip_clip_vit_b_16(seed=0)  # CLIP_VITB16 / SLIP_VITB16 geometry, full depth, trained-like values
    ref_enc = oracle.RefSlipVideoTextEncoder(copy.deepcopy(model))
    enc = B200SlipVideoTextEncoder(model.state_dict(), num_frames=4).to(dev)
    return ref_enc, enc


def test_slip_vit_b_16_full_depth_matches_oracle(full, dev):
    import oracle
    ref_enc, enc = full
    g = torch.Generator().manual_seed(99)
    video = torch.randn(5, 4, 3, 224, 224, generator=g)
    ids = torch.cat([oracle.tokenize_synthetic(8, (4, 40), seed=17), oracle.tokenize_synthetic(2, 77, seed=18)])
    with torch.inference_mode():
        ev, et = ref_enc.encode_video(video), ref_enc.encode_text({"input_ids": ids.long()})
        gv, gt = enc.encode_video(video.to(dev)).cpu(), enc.encode_text({"input_ids": ids.to(dev)}).cpu()
    cos, max_abs, centred = _report("slip video", gv, ev)
    assert cos >= 0.9995 and max_abs <= 5e-3 and centred <= 0.1
    cos, max_abs, centred = _report("slip text", gt, et)
    assert cos >= 0.9995 and max_abs <= 5e-3 and centred <= 0.1


@pytest.mark.parametrize("depth", [1, 4])
def test_slip_truncated_towers_match_oracle(depth, dev):
    """Shallow towers localise an error: depth 1 is the embedding (bias fold, no ln_pre: statistics of the raw tokens),
    one block with the erf-GELU epilogue, and the final norm + image_projection."""
    import oracle
    from fitclip_b200 import B200SlipVideoTextEncoder
    model = oracle.slip_clip_vit_b_16(seed=3, vision_layers=depth, transformer_layers=1)
    ref_enc = oracle.RefSlipVideoTextEncoder(copy.deepcopy(model))
    enc = B200SlipVideoTextEncoder(model.state_dict(), num_frames=2).to(dev)
    video = torch.randn(4, 2, 3, 224, 224, generator=torch.Generator().manual_seed(depth))
    with torch.inference_mode():
        expect = ref_enc.encode_video(video)
        got = enc.encode_video(video.to(dev)).cpu()
    cos, max_abs, centred = _report(f"slip depth {depth}", got, expect)
    assert cos >= 0.9995 and max_abs <= 5e-3 and centred <= 0.1


def test_slip_vit_l_16_geometry_runs_and_matches(dev):
    """``CLIP_VITL16`` / ``SLIP_VITL16`` (slip.py:618-640): 1024 wide, 16 heads; two layers keep the CPU oracle fast."""
    import oracle
    from fitclip_b200 import B200SlipVideoTextEncoder
    model = oracle.slip_clip_vit_b_16(seed=5, vision_width=1024, vision_layers=2, transformer_layers=2)
    ref_enc = oracle.RefSlipVideoTextEncoder(copy.deepcopy(model))
    enc = B200SlipVideoTextEncoder(model.state_dict(), num_frames=2).to(dev)
    assert enc.model.config["vision_width"] == 1024
    video = torch.randn(3, 2, 3, 224, 224, generator=torch.Generator().manual_seed(8))
    with torch.inference_mode():
        expect = ref_enc.encode_video(video)
        got = enc.encode_video(video.to(dev)).cpu()
    cos, max_abs, centred = _report("slip vit-l/16", got, expect)
    assert cos >= 0.9995 and max_abs <= 5e-3 and centred <= 0.1


@pytest.mark.parametrize("layers", [2, 12])
def test_slip_vit_s_16_narrow_heads_match_oracle(layers, dev):
    """``CLIP_VITS16`` / ``SLIP_VITS16`` (slip.py:566-590): 384 wide, 12 heads of 32 -- run as 12 zero-padded 64-wide heads
    (``fc_config.vision_attn_width = 768``); 12 layers is the full depth of the reference's model."""
    import oracle
    from fitclip_b200 import B200SlipVideoTextEncoder, load_slip_model
    model = oracle.slip_clip_vit_b_16(seed=6, vision_width=384, vision_heads=12, vision_layers=layers, transformer_layers=2)
    ref_enc = oracle.RefSlipVideoTextEncoder(copy.deepcopy(model))
    import argparse
    ckpt = {"args": argparse.Namespace(model="SLIP_VITS16"), "state_dict": {"module." + k: v for k, v in model.state_dict().items()}}
    enc = B200SlipVideoTextEncoder(load_slip_model(ckpt), num_frames=2).to(dev)
    assert enc.model.vision_heads == 12 and enc.model.config["vision_attn_width"] == 768
    video = torch.randn(4, 2, 3, 224, 224, generator=torch.Generator().manual_seed(layers))
    with torch.inference_mode():
        expect = ref_enc.encode_video(video)
        got = enc.encode_video(video.to(dev)).cpu()
    cos, max_abs, centred = _report(f"slip vit-s/16 x{layers}", got, expect)
    assert cos >= 0.9995 and max_abs <= 5e-3 and centred <= 0.1


def test_engine_rejects_ln_pre_for_the_timm_tower(ref, dev):
    """FC_TOWER_TIMM has no ln_pre slot: handing it one is an error, not a silently ignored tensor."""
    from fitclip_b200 import B200SlipClip, _lib
    model = B200SlipClip(ref["checkpoint_1"]).to(dev)
    handle = model._native(dev)
    w = torch.ones(64, device=dev)
    rc = _lib.load().fc_model_set_param(handle, b"visual.ln_pre.weight", w.data_ptr(), 64, _lib.stream_ptr(dev))
    assert rc != 0
