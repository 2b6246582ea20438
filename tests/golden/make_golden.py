"""Generates tests/golden/tiny_clip.pt from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference holds no golden vectors for this path, and its own CLIP dependency cannot be imported
here, so these fixtures freeze the restated oracle on a tiny CLIP with the real structure (ViT patch tower + causal
text tower, head dim 64).  tests/test_golden.py checks (a) the oracle still reproduces them (CPU) and (b) the CUDA path
matches them (GPU)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.encoder_ref import ref_batch_scores, ref_retrieval_scores  # noqa: E402

TINY = dict(embed_dim=64, image_resolution=64, vision_layers=2, vision_width=64, vision_patch_size=16,
            context_length=77, vocab_size=512, transformer_width=64, transformer_heads=1, transformer_layers=2)


def main() -> None:
    torch.set_num_threads(1)
    m1 = oracle.clip_vit_b_16(seed=0, **TINY)
    m2 = oracle.clip_vit_b_16(seed=1, **TINY)
    g = torch.Generator().manual_seed(1234)
    video = torch.randn(12, 2, 3, 64, 64, generator=g)
    ids = oracle.tokenize_synthetic(12, (4, 77), seed=4321, vocab_size=512)
    sd1 = {k: v.clone() for k, v in m1.state_dict().items()}
    sd2 = {k: v.clone() for k, v in m2.state_dict().items()}
    enc = oracle.RefClipVideoTextEncoder(m1)
    enc2 = oracle.RefClipVideoTextEncoder(m2)
    with torch.inference_mode():
        v, t = enc(video, {"input_ids": ids})
        scores = ref_retrieval_scores(t, v)
        metrics = oracle.ref_retrieval_metrics(scores)
        batch_scores = ref_batch_scores(v, t, 1 / 0.015)
        loss = oracle.ref_nce_loss(batch_scores)
        w = oracle.ref_wise(enc, enc2, weight_for_2=0.4)
        wv, wt = w(video, {"input_ids": ids})
        teacher_scores = ref_batch_scores(*enc2(video, {"input_ids": ids}), 1 / 0.015)
        ts_loss = oracle.ref_teacher_student_nce_loss(batch_scores, teacher_scores, reduction="batchmean")
    out = {
        "config": TINY, "state_dict_1": sd1, "state_dict_2": sd2, "video": video, "input_ids": ids,
        "video_emb": v.clone(), "text_emb": t.clone(), "scores": scores.clone(),
        "ranks": metrics["rank"].clone(), "r1": metrics["r1"].clone(), "r5": metrics["r5"].clone(), "r10": metrics["r10"].clone(),
        "mr": metrics["mr"].clone(), "loss_val": loss.clone(), "ts_loss": ts_loss.clone(),
        "wise_0.4_video_emb": wv.clone(), "wise_0.4_text_emb": wt.clone(),
        "wise_0.4_text_projection": dict(w.named_parameters())["model.text_projection"].detach().clone(),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_clip.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; ranks", out["ranks"].tolist(), "mr", int(out["mr"]))


if __name__ == "__main__":
    main()
