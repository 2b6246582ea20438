"""Pins the oracle's SLIP-layout restatement (oracle/slip_ref.py), on CPU:

* timm's ``VisionTransformer`` (third-party, not installed, not vendored by the reference) against an INDEPENDENT
  implementation of the same architecture in the image, ``transformers.ViTModel`` with ``layer_norm_eps=1e-6``,
  ``hidden_act="gelu"``, ``qkv_bias=True``, no pooler;
* ``slip.CLIP.encode_image / encode_text`` and ``SlipVideoTextEncoder.encode_video / encode_text`` against the outputs of
  the reference's own classes run in the build container (tests/golden/make_reference_slip_golden.py);
* the host-side name mapping of ``B200SlipClip`` (no GPU: only the list handed to the engine is inspected)."""
import os

import pytest
import torch

import oracle
from oracle.slip_ref import TimmVisionTransformer, perturb_timm_trained_like

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_slip.pt")


@pytest.fixture(scope="module")
def ref():
    return torch.load(GOLDEN, map_location="cpu", weights_only=False)


def _hf_vit_from_timm(vit: TimmVisionTransformer, img_size: int, patch: int, heads: int):
    from transformers import ViTConfig, ViTModel
    W, depth = vit.embed_dim, len(vit.blocks)
    hf = ViTModel(ViTConfig(hidden_size=W, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * W,
                            hidden_act="gelu", layer_norm_eps=1e-6, image_size=img_size, patch_size=patch, qkv_bias=True,
                            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
                  add_pooling_layer=False).eval()
    sd = vit.state_dict()
    new = {"embeddings.cls_token": sd["cls_token"], "embeddings.position_embeddings": sd["pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
           "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(depth):
        t, h = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, qb = sd[t + "attn.qkv.weight"], sd[t + "attn.qkv.bias"]
        for j, n in enumerate(("query", "key", "value")):  # timm packs q, k, v along the output dimension
            new[h + f"attention.attention.{n}.weight"] = qw[j * W:(j + 1) * W]
            new[h + f"attention.attention.{n}.bias"] = qb[j * W:(j + 1) * W]
        for a, b in (("attn.proj", "attention.output.dense"), ("norm1", "layernorm_before"), ("norm2", "layernorm_after"),
                     ("mlp.fc1", "intermediate.dense"), ("mlp.fc2", "output.dense")):
            new[h + b + ".weight"], new[h + b + ".bias"] = sd[t + a + ".weight"], sd[t + a + ".bias"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return hf


@pytest.mark.parametrize("img,patch,width,depth,heads", [(32, 16, 64, 2, 1), (48, 16, 128, 3, 2), (28, 14, 192, 2, 3)])
def test_timm_vit_restatement_matches_hf_vit(img, patch, width, depth, heads):
    torch.manual_seed(3)
    vit = TimmVisionTransformer(img, patch, width, depth, heads).eval()
    perturb_timm_trained_like(vit, seed=11)
    hf = _hf_vit_from_timm(vit, img, patch, heads)
    x = torch.randn(5, 3, img, img)
    with torch.inference_mode():
        got = vit(x)
        expect = hf(pixel_values=x).last_hidden_state[:, 0]
    assert got.shape == (5, width)
    assert torch.allclose(got, expect, atol=2e-5, rtol=1e-5), (got - expect).abs().max()


def _oracle_from(ref, which: str):
    model = oracle.slip_clip_vit_b_16(seed=123, trained_like=False, **ref["config"])
    model.load_state_dict({k.replace("module.", ""): v for k, v in ref[which].items()})
    return model.eval()


def test_oracle_slip_clip_matches_reference_classes(ref):
    model = _oracle_from(ref, "checkpoint_1")
    with torch.inference_mode():
        img = model.encode_image(ref["video"][:, 0])
        txt = model.encode_text(ref["input_ids"].long())
    assert torch.allclose(img, ref["image_features"], atol=1e-6, rtol=1e-6)
    assert torch.allclose(txt, ref["text_features"], atol=2e-6, rtol=1e-6)


def test_oracle_slip_wrapper_matches_reference_wrapper(ref):
    enc = oracle.RefSlipVideoTextEncoder(_oracle_from(ref, "checkpoint_1"), num_frames=3)
    with torch.inference_mode():
        v = enc.encode_video(ref["video"])
        t = enc.encode_text({"input_ids": ref["input_ids"].long()})
    assert torch.allclose(v, ref["wrapper_video_emb"], atol=1e-6, rtol=0)
    assert torch.allclose(t, ref["wrapper_text_emb"], atol=1e-6, rtol=0)


def test_slip_layout_keeps_checkpoint_names_and_maps_them_for_the_engine(ref):
    from fitclip_b200 import B200SlipClip, B200SlipVideoTextEncoder
    from fitclip_b200.encoder import infer_config
    model = B200SlipClip(ref["checkpoint_1"])  # DDP-prefixed names, as load_model receives them
    own = [n for n, _ in model.named_parameters()]
    assert sorted(own) == sorted(k.replace("module.", "") for k in ref["checkpoint_1"])
    enc = B200SlipVideoTextEncoder(model, num_frames=3)
    assert sorted(n for n, _ in enc.named_parameters()) == sorted(ref["wrapper_param_names"])  # logit_scale dropped
    assert model.config["vision_tower"] == 1 and model.config["image_resolution"] == 32
    # what the engine is handed: the OpenAI names of the same geometry, minus ln_pre (the timm tower has none)
    handed = dict(model._engine_params())
    tiny = {k: v for k, v in ref["config"].items()}
    openai = oracle.clip_vit_b_16(seed=0, embed_dim=tiny["embed_dim"], image_resolution=tiny["img_size"],
                                  vision_layers=tiny["vision_layers"], vision_width=tiny["vision_width"],
                                  vision_patch_size=tiny["patch_size"], context_length=tiny["context_length"],
                                  vocab_size=tiny["vocab_size"], transformer_width=tiny["transformer_width"],
                                  transformer_heads=tiny["transformer_heads"],
                                  transformer_layers=tiny["transformer_layers"]).state_dict()
    expect = {k: v for k, v in openai.items() if "ln_pre" not in k and k != "logit_scale"}
    assert sorted(handed) == sorted(expect)
    for k, v in handed.items():
        assert v.numel() == expect[k].numel(), k
    assert {k: v for k, v in infer_config(openai).items()} == {k: v for k, v in model.config.items() if k != "vision_tower"}
    # the patch-embedding bias rides on the positional rows of the patch tokens, not on the class token's row
    sd = {k.replace("module.", ""): v for k, v in ref["checkpoint_1"].items()}
    pos = handed["visual.positional_embedding"]
    assert torch.equal(pos[0], sd["visual.pos_embed"][0, 0])
    assert torch.allclose(pos[1:], sd["visual.pos_embed"][0, 1:] + sd["visual.patch_embed.proj.bias"], atol=0, rtol=0)
    assert torch.equal(handed["visual.class_embedding"], sd["visual.cls_token"].reshape(-1))
    assert torch.equal(handed["visual.proj"], sd["image_projection"])


def test_slip_vit_small_heads_are_padded_to_64_wide_slots():
    """``vit_small_mocov3_patch16_224`` (slip.py:566-569): 384 wide, 12 heads of 32.  The engine's attention kernels take
    64-wide heads, so every head gets a 64-wide slot of zero-padded q / k / v rows (q scaled by sqrt 2: the kernels' softmax
    scale is 1/8) and zero out_proj columns -- checked here in fp32 on the tensors handed to the engine."""
    from fitclip_b200 import B200SlipClip, _lib
    sd = oracle.slip_clip_vit_b_16(seed=0, img_size=32, patch_size=16, vision_width=384, vision_layers=1, vision_heads=12,
                                   embed_dim=64, context_length=16, vocab_size=512, transformer_width=64,
                                   transformer_heads=1, transformer_layers=1).state_dict()
    with pytest.raises(_lib.FitclipError, match="vision_heads"):
        B200SlipClip(sd)  # 6 heads of 64 (timm's stock ViT-S) or 12 of 32 (SLIP's): the state dict cannot tell
    with pytest.raises(_lib.FitclipError, match="divide 64"):
        B200SlipClip(sd, vision_heads=8)  # head dimension 48
    assert "vision_attn_width" not in B200SlipClip(sd, vision_heads=6).config
    model = B200SlipClip(sd, vision_heads=12)
    assert model.config["vision_attn_width"] == 768 and model.config["vision_width"] == 384
    handed = dict(model._engine_params())
    w2, b2 = handed["visual.transformer.resblocks.0.attn.in_proj_weight"], handed["visual.transformer.resblocks.0.attn.in_proj_bias"]
    p2 = handed["visual.transformer.resblocks.0.attn.out_proj.weight"]
    assert w2.shape == (3 * 768, 384) and b2.shape == (3 * 768,) and p2.shape == (384, 768)
    x = torch.randn(7, 384, generator=torch.Generator().manual_seed(1))
    qkv = (x @ sd["visual.blocks.0.attn.qkv.weight"].T + sd["visual.blocks.0.attn.qkv.bias"]).view(7, 3, 12, 32)
    att = torch.softmax(torch.einsum("ihd,jhd->hij", qkv[:, 0], qkv[:, 1]) * 32 ** -0.5, -1)
    expect = torch.einsum("hij,jhd->ihd", att, qkv[:, 2]).reshape(7, 384) @ sd["visual.blocks.0.attn.proj.weight"].T
    qkv2 = (x @ w2.T + b2).view(7, 3, 12, 64)
    att2 = torch.softmax(torch.einsum("ihd,jhd->hij", qkv2[:, 0], qkv2[:, 1]) * 0.125, -1)  # what the kernels compute
    got = torch.einsum("hij,jhd->ihd", att2, qkv2[:, 2]).reshape(7, 768) @ p2.T
    assert torch.allclose(got, expect, atol=1e-6, rtol=1e-5)


def test_slip_layout_student_is_refused_by_the_trainer_but_fine_as_teacher(ref):
    from fitclip_b200 import B200SlipVideoTextEncoder
    from fitclip_b200.training import ClipTrainer
    enc = B200SlipVideoTextEncoder(ref["checkpoint_1"], num_frames=3)
    with pytest.raises(NotImplementedError, match="SLIP-layout"):
        ClipTrainer(enc.model, kernels=object())
