"""Outputs of the REFERENCE'S OWN in-tree code (tests/golden/reference_outputs.pt, produced by
tests/golden/make_reference_golden.py in the build container, where /root/reference is mounted) against

* the CPU oracle (`-m "not gpu"`): pins the restatement to the reference for every function whose code is in the
  reference tree -- the ClipVideoTextEncoder wrapper, the text tower twin in slip.py, WiSE, the losses, Rank/MedianRank,
  the eval frame sampler;
* the CUDA path through the C ABI (`-m gpu`).

Tolerances: fp32 CPU vs fp32 CPU -> 1e-6 (same formulas, possibly different op order); WiSE and ranks bit-exact;
bf16 tensor-core path vs fp32 reference -> cosine >= 0.999 per vector (BASELINE.json) and max-abs <= 2e-2."""
import os

import pytest
import torch
import torch.nn.functional as F

import oracle

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.pt")


@pytest.fixture(scope="module")
def ref():
    return torch.load(PATH, map_location="cpu", weights_only=False)


def _oracle_encoder(ref, which="state_dict_1"):
    model = oracle.CLIP(**ref["config"]).float().eval()
    model.load_state_dict(ref[which])
    return oracle.RefClipVideoTextEncoder(model, num_frames=3)


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_wrapper_matches_reference_wrapper(ref):
    enc = _oracle_encoder(ref)
    with torch.inference_mode():
        v, t = enc(ref["video"], {"input_ids": ref["input_ids"]})
    assert torch.allclose(v, ref["wrapper_video_emb"], atol=1e-6, rtol=0)
    assert torch.allclose(t, ref["wrapper_text_emb"], atol=1e-6, rtol=0)
    # per-frame normalisation then a plain mean: pooled video vectors are NOT unit length (clip_video_text_encoder.py:85-89)
    assert (ref["wrapper_video_emb"].norm(dim=-1) < 1 - 1e-4).any()
    assert [n for n, _ in enc.named_parameters()] == ref["wrapper_param_names"]


def test_oracle_text_tower_matches_slip_twin(ref):
    """aligner/encoder/slip.py:350-480 run on the oracle's text weights: LayerNorm, QuickGELU, the residual block, the
    causal mask and the EOT pooling of the restatement agree with the reference's own classes."""
    enc = _oracle_encoder(ref)
    with torch.inference_mode():
        feats = enc.model.encode_text(ref["input_ids"])
    assert torch.allclose(feats, ref["slip_text_features"], atol=2e-6, rtol=1e-6), \
        (feats - ref["slip_text_features"]).abs().max().item()


@pytest.mark.parametrize("w", [0.4, 0.5])
def test_oracle_wise_matches_reference(ref, w):
    a, b = _oracle_encoder(ref, "state_dict_1"), _oracle_encoder(ref, "state_dict_2")
    ours = oracle.ref_wise_state_dict(a, b, weight_for_2=w)
    for k, expect in ref[f"wise_{w}_state_dict"].items():
        assert torch.equal(ours[k], expect), k
    if w == 0.4:
        merged = oracle.ref_wise(a, b, weight_for_2=0.4)
        with torch.inference_mode():
            v, t = merged(ref["video"], {"input_ids": ref["input_ids"]})
        assert torch.allclose(v, ref["wise_video_emb"], atol=1e-6, rtol=0)
        assert torch.allclose(t, ref["wise_text_emb"], atol=1e-6, rtol=0)


def test_oracle_losses_match_reference(ref):
    s, t = ref["loss_scores"], ref["loss_teacher_scores"]
    for red in ("mean", "sum", "none"):
        assert torch.allclose(oracle.ref_nce_loss(s, reduction=red), ref[f"nce_{red}"], atol=1e-6)
    for red in ("sum", "batchmean"):
        assert torch.allclose(oracle.ref_teacher_student_nce_loss(s, t, reduction=red), ref[f"ts_nce_{red}"], atol=1e-5,
                              rtol=1e-6)
    assert torch.allclose(ref["nce_module"], ref["nce_mean"])
    assert torch.allclose(ref["ts_nce_module"], ref["ts_nce_batchmean"])


def test_oracle_ranks_match_reference(ref):
    for case in ref["rank_cases"]:
        ranks = oracle.ref_rank(case["scores"], case["target"])
        assert torch.equal(ranks, case["ranks"])
        assert torch.equal(oracle.ref_stable_rank(case["scores"], case["target"]), case["ranks"])  # tie-free matrices
        assert torch.equal(oracle.ref_median_rank(ranks), case["median_rank"])


def test_frame_sampler_matches_reference(ref):
    from fitclip_b200.frame_sampler import UniformFrameSampler
    for case in ref["uniform_sampler_cases"]:
        got = [int(i) for i in UniformFrameSampler(case["max_frames"])(case["start"], case["end"], 30.0)]
        assert got == case["indices"], case


# ------------------------------------------------------------------------------------------------ GPU: CUDA path vs reference
@pytest.mark.gpu
def test_cuda_encoder_matches_reference_wrapper(ref, dev):
    from fitclip_b200 import B200ClipVideoTextEncoder
    enc = B200ClipVideoTextEncoder(ref["state_dict_1"], num_frames=3).to(dev)
    with torch.inference_mode():
        v, t = enc(ref["video"].to(dev), {"input_ids": ref["input_ids"].to(dev)})
    for got, expect in ((v.cpu(), ref["wrapper_video_emb"]), (t.cpu(), ref["wrapper_text_emb"])):
        assert F.cosine_similarity(got, expect).min().item() >= 0.999
        assert (got - expect).abs().max().item() <= 2e-2


@pytest.mark.gpu
def test_cuda_wise_matches_reference(ref, dev):
    from fitclip_b200 import B200ClipVideoTextEncoder
    from fitclip_b200.wise import wise, wise_state_dict
    a = B200ClipVideoTextEncoder(ref["state_dict_1"], num_frames=3).to(dev)
    b = B200ClipVideoTextEncoder(ref["state_dict_2"], num_frames=3).to(dev)
    for w in (0.4, 0.5):
        ours = wise_state_dict(a, b, weight_for_2=w)
        for k, expect in ref[f"wise_{w}_state_dict"].items():
            assert torch.equal(ours[k].cpu(), expect), (w, k)  # bit-exact: two rounded products + a rounded sum
    merged = wise(a, b, weight_for_2=0.4)
    with torch.inference_mode():
        v, t = merged(ref["video"].to(dev), {"input_ids": ref["input_ids"].to(dev)})
    assert F.cosine_similarity(v.cpu(), ref["wise_video_emb"]).min().item() >= 0.999
    assert F.cosine_similarity(t.cpu(), ref["wise_text_emb"]).min().item() >= 0.999


@pytest.mark.gpu
def test_cuda_losses_match_reference(ref, dev):
    from fitclip_b200 import ops
    s, t = ref["loss_scores"].to(dev), ref["loss_teacher_scores"].to(dev)
    got = ops.nce_loss(s).item()
    assert abs(got - ref["nce_mean"].item()) <= 1e-4 * max(1.0, abs(ref["nce_mean"].item()))
    got = ops.teacher_student_nce_loss(s, t).item()
    assert abs(got - ref["ts_nce_batchmean"].item()) <= 1e-4 * max(1.0, abs(ref["ts_nce_batchmean"].item()))


@pytest.mark.gpu
def test_cuda_ranks_match_reference(ref, dev):
    from fitclip_b200 import ops
    from fitclip_b200.metrics import MedianRank, Rank
    for case in ref["rank_cases"]:
        scores, target = case["scores"].to(dev), case["target"].to(dev)
        assert torch.equal(ops.rank_from_scores(scores, target).cpu(), case["ranks"])
        rank_m, med_m = Rank(), MedianRank()
        half = scores.shape[0] // 2
        for lo, hi in ((0, half), (half, scores.shape[0])):  # two update() calls like a validation epoch
            rank_m.update(scores[lo:hi], target[lo:hi])
            med_m.update(scores[lo:hi], target[lo:hi])
        assert torch.equal(rank_m.compute().cpu(), case["ranks"])
        assert int(med_m.compute()) == int(case["median_rank"])
