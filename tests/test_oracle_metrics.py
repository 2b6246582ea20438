"""Oracle self-consistency for the integer / scalar parts of the path (CPU)."""
import copy

import pytest
import torch

import oracle
from oracle.encoder_ref import ref_batch_scores, ref_class_embeddings, ref_retrieval_scores


def test_ref_rank_equals_bruteforce_count_on_tie_free_rows():
    torch.manual_seed(0)
    s = torch.randn(200, 77)
    target = torch.randint(0, 77, (200,))
    brute = torch.tensor([(s[i] > s[i, target[i]]).sum() for i in range(200)])
    assert torch.equal(oracle.ref_rank(s, target), brute)
    assert torch.equal(oracle.ref_stable_rank(s, target), brute)


def test_stable_rank_is_stable_sort_rank_under_ties():
    torch.manual_seed(1)
    s = torch.randint(0, 3, (100, 40)).float()
    target = torch.randint(0, 40, (100,))
    stable = torch.where(s.argsort(dim=1, descending=True, stable=True) == target.unsqueeze(-1))[1]
    assert torch.equal(oracle.ref_stable_rank(s, target), stable)


def test_median_rank_is_lower_median_plus_one():
    assert int(oracle.ref_median_rank(torch.tensor([0, 1, 2, 3]))) == 2   # torch.median -> lower middle (1) + 1
    assert int(oracle.ref_median_rank(torch.tensor([5, 0, 9]))) == 6
    assert oracle.ref_median_rank(torch.tensor([3])).dtype == torch.int64


def test_recall_is_mean_of_rank_below_k():
    torch.manual_seed(2)
    s = torch.randn(300, 50)
    target = torch.randint(0, 50, (300,))
    ranks = oracle.ref_rank(s, target)
    for k in (1, 5, 10):
        assert float(oracle.ref_recall_at_k(s, target, k)) == float((ranks < k).float().mean())
    with pytest.raises(RuntimeError):
        oracle.ref_recall_at_k(s[:, :5], target.clamp(max=4), 10)  # k must stay below the candidate count


def test_retrieval_metrics_on_a_known_matrix():
    s = torch.eye(12) + 0.01 * torch.arange(12).float().unsqueeze(0)  # diagonal wins everywhere
    m = oracle.ref_retrieval_metrics(s)
    assert float(m["r1"]) == 1.0 and int(m["mr"]) == 1 and m["rank"].tolist() == [0] * 12


def test_wise_is_the_literal_expression_and_leaves_inputs_alone():
    tiny = dict(embed_dim=64, image_resolution=32, vision_layers=1, vision_width=64, vision_patch_size=16,
                context_length=77, vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=1)
    e1 = oracle.RefClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **tiny))
    e2 = oracle.RefClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1, **tiny))
    before = copy.deepcopy(e1.state_dict())
    w = oracle.ref_wise(e1, e2, weight_for_2=0.4)
    for k, v in w.state_dict().items():
        assert torch.equal(v, (1 - 0.4) * e1.state_dict()[k] + 0.4 * e2.state_dict()[k])
        assert torch.equal(e1.state_dict()[k], before[k])
    assert "model.logit_scale" not in w.state_dict()  # removed by the wrapper (clip_video_text_encoder.py:75-77)


def test_encode_video_normalises_per_frame_then_means_without_renormalising():
    tiny = dict(embed_dim=64, image_resolution=32, vision_layers=1, vision_width=64, vision_patch_size=16,
                context_length=77, vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=1)
    enc = oracle.RefClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **tiny))
    video = torch.randn(3, 4, 3, 32, 32)
    with torch.inference_mode():
        out = enc.encode_video(video)
        frames = enc.model.encode_image(video.view(-1, 3, 32, 32))
    expect = torch.nn.functional.normalize(frames, dim=-1).view(3, 4, -1).mean(1)
    assert torch.allclose(out, expect, atol=1e-6)
    assert (out.norm(dim=-1) < 1.0).all()  # a mean of unit vectors is shorter than 1: no renormalisation happened


def test_scores_orientation_and_scale_precedence():
    v, t = torch.randn(4, 8), torch.randn(5, 8)
    assert ref_retrieval_scores(t, v).shape == (5, 4)  # rows = texts, columns = videos
    assert torch.allclose(ref_batch_scores(v, v, 66.667), (66.667 * v) @ v.T)


def test_losses_match_their_definitions():
    torch.manual_seed(3)
    s, t = torch.randn(6, 6) * 3, torch.randn(6, 6) * 3
    lsm = torch.log_softmax
    assert torch.allclose(oracle.ref_nce_loss(s), -(lsm(s, -1).diag().mean() + lsm(s.T, -1).diag().mean()))
    pt, pc = torch.softmax(t, -1), torch.softmax(t.T, -1)
    expect = ((pt * (pt.log() - lsm(s, -1))).sum() + (pc * (pc.log() - lsm(s.T, -1))).sum()) / 6
    assert torch.allclose(oracle.ref_teacher_student_nce_loss(s, t, reduction="batchmean"), expect, atol=1e-6)


def test_class_embeddings_mean_over_templates():
    tiny = dict(embed_dim=64, image_resolution=32, vision_layers=1, vision_width=64, vision_patch_size=16,
                context_length=77, vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=1)
    enc = oracle.RefClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **tiny))
    ids = oracle.tokenize_synthetic(6 * 3, (3, 20), seed=5, vocab_size=64)
    with torch.inference_mode():
        labels = ref_class_embeddings(enc, ids, template_count=3, batch_size=4)
        direct = enc.encode_text({"input_ids": ids}).reshape(6, 3, -1).mean(1)
    assert labels.shape == (6, 64) and torch.allclose(labels, direct, atol=1e-6)


def test_tokenize_synthetic_layout():
    ids = oracle.tokenize_synthetic(50, (3, 77), seed=0)
    assert ids.dtype == torch.int32 and ids.shape == (50, 77)
    assert (ids[:, 0] == 49406).all()
    eot = ids.argmax(dim=-1)
    for i in range(50):
        assert ids[i, eot[i]] == 49407 and (ids[i, eot[i] + 1:] == 0).all() and (ids[i, 1:eot[i]] < 49406).all()
