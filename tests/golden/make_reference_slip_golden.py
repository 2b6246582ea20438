"""Golden vectors for the SLIP-layout models, produced by the REFERENCE'S OWN in-tree code run in the
build container (``python tests/golden/make_reference_slip_golden.py`` -> ``tests/golden/reference_slip.pt``).

What runs is the reference's ``aligner/encoder/slip.py`` class ``CLIP`` (``:399-480``: ``encode_image`` = vision model +
``image_projection``, ``encode_text``) and ``aligner/encoder/slip_video_text_encoder.py`` ``SlipVideoTextEncoder``
(``encode_video`` / ``encode_text`` / ``forward``, ``load_model``'s "module." handling through ``load_state_dict``), with the
third-party stubs of ``make_reference_golden.install_stubs`` and two stand-ins more: ``SimpleTokenizer`` (its vocabulary
file is not on disk) is replaced by a no-argument stub before the wrapper is constructed, and the vision model handed to
``slip.CLIP`` is the oracle's restatement of timm's ``VisionTransformer`` (timm is not installed; that restatement is
pinned separately against ``transformers.ViTModel``, tests/test_oracle_slip.py).  No reference file is copied or edited.

Also stores WiSE (``aligner/wise.py``) of two SLIP-layout encoders, which the product must reproduce name by name."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from make_reference_golden import REFERENCE, install_stubs  # noqa: E402

TINY = dict(img_size=32, patch_size=16, vision_width=64, vision_layers=2, vision_heads=1, embed_dim=64, context_length=16,
            vocab_size=512, transformer_width=64, transformer_heads=1, transformer_layers=2)


def main() -> None:
    assert os.path.isdir(REFERENCE), f"{REFERENCE} is not mounted: this script only runs in the build container"
    torch.set_num_threads(1)
    install_stubs()
    sys.path.insert(0, REFERENCE)
    from aligner import wise as ref_wise  # noqa: E402  (reference modules)
    from aligner.encoder import slip as ref_slip  # noqa: E402
    from aligner.encoder import slip_video_text_encoder as ref_wrapper  # noqa: E402

    ref_wrapper.SimpleTokenizer = lambda: None  # the BPE vocabulary file is not on disk; tokenisation is not on this path

    def reference_model(seed: int):
        src = oracle.slip_clip_vit_b_16(seed=seed, **TINY)
        vit = oracle.TimmVisionTransformer(TINY["img_size"], TINY["patch_size"], TINY["vision_width"],
                                           TINY["vision_layers"], TINY["vision_heads"])
        model = ref_slip.CLIP(embed_dim=64, vision_width=64, vision_model=vit, context_length=16, vocab_size=512,
                              transformer_width=64, transformer_heads=1, transformer_layers=2)
        # through the checkpoint route of load_model (slip_video_text_encoder.py:19-23): DDP-prefixed names
        ckpt = {"module." + k: v for k, v in src.state_dict().items()}
        model.load_state_dict({k.replace("module.", ""): v for k, v in ckpt.items()})
        return model.eval(), ckpt

    out = {"config": TINY, "reference_files": ["aligner/encoder/slip.py", "aligner/encoder/slip_video_text_encoder.py",
                                               "aligner/wise.py"]}
    g = torch.Generator().manual_seed(20221119)
    m1, ckpt1 = reference_model(0)
    m2, ckpt2 = reference_model(1)
    out["checkpoint_1"], out["checkpoint_2"] = ckpt1, ckpt2
    video = torch.randn(10, 3, 3, 32, 32, generator=g)
    ids = oracle.tokenize_synthetic(10, (3, 16), seed=78, context_length=16, vocab_size=512)
    with torch.inference_mode():
        out["image_features"] = m1.encode_image(video[:, 0]).clone()       # slip.py:462-466, un-normalised
        out["text_features"] = m1.encode_text(ids.long()).clone()          # slip.py:468-480
    enc1 = ref_wrapper.SlipVideoTextEncoder(m1, num_frames=3)
    enc2 = ref_wrapper.SlipVideoTextEncoder(m2, num_frames=3)
    assert not hasattr(enc1.model, "logit_scale")  # :33-35
    with torch.inference_mode():
        v, t = enc1(video=video, text={"input_ids": ids.long()})
    out.update(video=video, input_ids=ids, wrapper_video_emb=v.clone(), wrapper_text_emb=t.clone(),
               wrapper_param_names=[n for n, _ in enc1.named_parameters()])
    for w in (0.5, 0.4):
        with torch.inference_mode():
            wised = ref_wise.wise(enc1, enc2, weight_for_2=w)
            wv, wt = wised(video=video, text={"input_ids": ids.long()})
        out[f"wise_{w}_video_emb"], out[f"wise_{w}_text_emb"] = wv.clone(), wt.clone()
        sd = wised.state_dict()
        keep = ("model.visual.pos_embed", "model.visual.blocks.1.attn.qkv.weight", "model.image_projection",
                "model.visual.patch_embed.proj.bias", "model.transformer.resblocks.0.mlp.c_fc.weight")
        out[f"wise_{w}_state_dict"] = {k: sd[k].detach().clone() for k in keep}
    path = os.path.join(HERE, "reference_slip.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
