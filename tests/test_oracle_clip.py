"""Pins the oracle's restated CLIP (oracle/clip_ref.py) against an independent implementation that exists in this
image, ``transformers.CLIPModel``, under the documented weight mapping: the reference's own
CLIP dependency (openai/CLIP@b46f5ac) is not installed and the reference ships no golden vectors for this path."""
import pytest
import torch

import oracle

transformers = pytest.importorskip("transformers")

SMALL = dict(embed_dim=64, image_resolution=64, vision_layers=2, vision_width=128, vision_patch_size=16,
             context_length=77, vocab_size=1000, transformer_width=128, transformer_heads=2, transformer_layers=2)


def _hf_from_oracle(m: oracle.CLIP, cfg: dict):
    from transformers import CLIPConfig, CLIPModel
    config = CLIPConfig(
        text_config=dict(vocab_size=cfg["vocab_size"], hidden_size=cfg["transformer_width"],
                         intermediate_size=4 * cfg["transformer_width"], num_hidden_layers=cfg["transformer_layers"],
                         num_attention_heads=cfg["transformer_heads"], max_position_embeddings=cfg["context_length"],
                         hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2, bos_token_id=0, pad_token_id=1),
        vision_config=dict(hidden_size=cfg["vision_width"], intermediate_size=4 * cfg["vision_width"],
                           num_hidden_layers=cfg["vision_layers"], num_attention_heads=cfg["vision_width"] // 64,
                           image_size=cfg["image_resolution"], patch_size=cfg["vision_patch_size"],
                           hidden_act="quick_gelu", layer_norm_eps=1e-5),
        projection_dim=cfg["embed_dim"])
    hf = CLIPModel(config).eval().float()
    sd = m.state_dict()
    new = {}

    def blocks(src_prefix, dst_prefix, layers):
        for i in range(layers):
            s, d = f"{src_prefix}.resblocks.{i}.", f"{dst_prefix}.encoder.layers.{i}."
            wq, wk, wv = sd[s + "attn.in_proj_weight"].chunk(3)
            bq, bk, bv = sd[s + "attn.in_proj_bias"].chunk(3)
            for n, w, b in (("q_proj", wq, bq), ("k_proj", wk, bk), ("v_proj", wv, bv)):
                new[d + f"self_attn.{n}.weight"], new[d + f"self_attn.{n}.bias"] = w, b
            new[d + "self_attn.out_proj.weight"] = sd[s + "attn.out_proj.weight"]
            new[d + "self_attn.out_proj.bias"] = sd[s + "attn.out_proj.bias"]
            for a, b_ in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2")):
                new[d + b_ + ".weight"], new[d + b_ + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
            for a, b_ in (("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
                new[d + b_ + ".weight"], new[d + b_ + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]

    blocks("visual.transformer", "vision_model", cfg["vision_layers"])
    blocks("transformer", "text_model", cfg["transformer_layers"])
    new["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    new["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    new["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    new["vision_model.pre_layrnorm.weight"], new["vision_model.pre_layrnorm.bias"] = \
        sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"]
    new["vision_model.post_layernorm.weight"], new["vision_model.post_layernorm.bias"] = \
        sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]
    new["visual_projection.weight"] = sd["visual.proj"].T
    new["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    new["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    new["text_model.final_layer_norm.weight"], new["text_model.final_layer_norm.bias"] = \
        sd["ln_final.weight"], sd["ln_final.bias"]
    new["text_projection.weight"] = sd["text_projection"].T
    new["logit_scale"] = sd["logit_scale"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


@pytest.fixture(scope="module")
def pair():
    torch.manual_seed(0)
    m = oracle.clip_vit_b_16(seed=0, **SMALL)
    # make LayerNorm affine parameters non-trivial so the mapping is really exercised
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".ln_" in n or "ln_pre" in n or "ln_post" in n or "ln_final" in n:
                p.add_(0.1 * torch.randn_like(p))
    return m, _hf_from_oracle(m, SMALL)


def test_encode_image_matches_hf(pair):
    m, hf = pair
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    with torch.inference_mode():
        a = m.encode_image(x)
        b = hf.get_image_features(pixel_values=x)
    b = getattr(b, "pooler_output", b)
    assert (a - b).abs().max().item() <= 1e-4


def test_encode_text_matches_hf_including_ragged_eot(pair):
    m, hf = pair
    ids = oracle.tokenize_synthetic(5, (3, 77), seed=2, vocab_size=1000)
    with torch.inference_mode():
        a = m.encode_text(ids)
        b = hf.get_text_features(input_ids=ids.long())
    b = getattr(b, "pooler_output", b)
    assert (a - b).abs().max().item() <= 1e-4


def test_causal_mask_tokens_after_eot_do_not_matter(pair):
    m, _ = pair
    ids = oracle.tokenize_synthetic(2, 10, seed=3, vocab_size=1000)
    ids2 = ids.clone()
    ids2[:, 10:] = torch.randint(1, 900, ids2[:, 10:].shape, dtype=torch.int32)  # garbage after EOT (< EOT id)
    with torch.inference_mode():
        assert torch.allclose(m.encode_text(ids), m.encode_text(ids2), atol=1e-6)


def test_build_model_infers_geometry_and_names():
    m = oracle.clip_vit_b_16(seed=0, **SMALL)
    rebuilt = oracle.build_model(m.state_dict())
    assert [k for k, _ in rebuilt.named_parameters()] == [k for k, _ in m.named_parameters()]
    x = torch.randn(1, 3, 64, 64)
    with torch.inference_mode():
        assert torch.equal(rebuilt.encode_image(x), m.encode_image(x))


def test_vit_b_16_parameter_inventory():
    m = oracle.clip_vit_b_16(seed=0)
    names = [k for k, _ in m.named_parameters()]
    assert len(names) == 302 and "logit_scale" in names  # 301 + logit_scale
    assert sum(p.numel() for p in m.parameters()) == 149_620_737
    assert m.visual.positional_embedding.shape == (197, 768) and m.visual.proj.shape == (768, 512)
    assert m.transformer.resblocks[0].attn.in_proj_weight.shape == (1536, 512)
