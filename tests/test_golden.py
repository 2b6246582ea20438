"""Golden vectors frozen from the oracle (tests/golden/make_golden.py).  CPU: the oracle still reproduces them.
GPU: the CUDA path matches them (embeddings within the bf16 tolerance, WiSE bits exactly, ranks of the kernel's own
score matrix exactly)."""
import os

import pytest
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_clip.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, map_location="cpu")


def test_oracle_reproduces_golden(golden):
    import oracle
    from oracle.encoder_ref import ref_batch_scores, ref_retrieval_scores
    m = oracle.CLIP(**golden["config"]).float().eval()
    m.load_state_dict(golden["state_dict_1"])
    enc = oracle.RefClipVideoTextEncoder(m)
    with torch.inference_mode():
        v, t = enc(golden["video"], {"input_ids": golden["input_ids"]})
    assert torch.allclose(v, golden["video_emb"], atol=1e-5)
    assert torch.allclose(t, golden["text_emb"], atol=1e-5)
    # integer results are recomputed from the FROZEN scores, so they must match exactly on any machine
    met = oracle.ref_retrieval_metrics(golden["scores"])
    assert torch.equal(met["rank"], golden["ranks"])
    assert int(met["mr"]) == int(golden["mr"])
    for k in ("r1", "r5", "r10"):
        assert float(met[k]) == float(golden[k])
    assert torch.allclose(ref_retrieval_scores(t, v), golden["scores"], atol=1e-5)
    loss = oracle.ref_nce_loss(ref_batch_scores(golden["video_emb"], golden["text_emb"], 1 / 0.015))
    assert abs(float(loss) - float(golden["loss_val"])) <= 1e-4


def test_seeded_init_reproduces_golden_weights(golden):
    import oracle
    m = oracle.clip_vit_b_16(seed=0, **golden["config"])
    for k, v in m.state_dict().items():
        assert torch.equal(v, golden["state_dict_1"][k]), k


@pytest.mark.gpu
def test_cuda_path_matches_golden(golden, dev):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder, ops, retrieval_ranks, wise
    enc = B200ClipVideoTextEncoder(golden["state_dict_1"], num_frames=2).to(dev)
    v, t = enc(golden["video"].to(dev), {"input_ids": golden["input_ids"].to(dev)})
    assert F.cosine_similarity(v.cpu(), golden["video_emb"]).min().item() >= 0.999
    assert F.cosine_similarity(t.cpu(), golden["text_emb"]).min().item() >= 0.999
    assert (v.cpu() - golden["video_emb"]).abs().max().item() <= 2e-2
    assert (t.cpu() - golden["text_emb"]).abs().max().item() <= 2e-2
    # similarity + rank on the GOLDEN embeddings: split-bf16 scores within 2e-6 of fp32, ranks identical
    sim = ops.Similarity(golden["text_emb"].to(dev), golden["video_emb"].to(dev), 3)
    assert (sim.scores().cpu() - golden["scores"]).abs().max().item() <= 2e-6
    ranks = retrieval_ranks(golden["text_emb"].to(dev), golden["video_emb"].to(dev)).cpu()
    assert torch.equal(ranks, golden["ranks"])
    recall, median, _ = ops.metrics_from_ranks(ranks.to(dev), 12)
    assert int(median) == int(golden["mr"])
    assert [float(x) for x in recall] == [float(golden[k]) for k in ("r1", "r5", "r10")]
    # per-batch loss/val (temperature 0.015, config/trainer.yaml:19)
    scores = ops.Similarity(golden["video_emb"].to(dev), golden["text_emb"].to(dev), 3).scores(alpha=1 / 0.015)
    assert abs(float(ops.nce_loss(scores)) - float(golden["loss_val"])) <= 1e-3
    # WiSE w = 0.4: parameter bits exact, embeddings within tolerance
    enc2 = B200ClipVideoTextEncoder(golden["state_dict_2"], num_frames=2).to(dev)
    w = wise(enc, enc2, weight_for_2=0.4)
    assert torch.equal(dict(w.named_parameters())["model.text_projection"].cpu(), golden["wise_0.4_text_projection"])
    wv, wt = w(golden["video"].to(dev), {"input_ids": golden["input_ids"].to(dev)})
    assert F.cosine_similarity(wv.cpu(), golden["wise_0.4_video_emb"]).min().item() >= 0.999
    assert F.cosine_similarity(wt.cpu(), golden["wise_0.4_text_emb"]).min().item() >= 0.999
