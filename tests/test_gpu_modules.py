"""Task-level modules on the CUDA path against the oracle: retrieval evaluation (config 2 shape, scaled down),
zero-shot classification (config 3 shape, scaled down) and teacher-student scoring (config 5 shape, scaled down)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(vision_layers=2, transformer_layers=2)  # real widths / heads / sequence lengths, 2 layers per tower


@pytest.fixture(scope="module")
def pair(dev):
    import oracle
    from fitclip_b200 import B200ClipVideoTextEncoder
    m = oracle.clip_vit_b_16(seed=0, **TINY)
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(m))
    enc = B200ClipVideoTextEncoder(m.state_dict(), num_frames=2).to(dev)
    return ref, enc


def test_retrieval_module_matches_reference_flow(pair, dev):
    import oracle
    from oracle.encoder_ref import ref_batch_scores
    from fitclip_b200 import TextVideoRetrievalModule, ops
    ref, enc = pair
    module = TextVideoRetrievalModule(enc, init_temperature=0.015, fit_temperature=False, compute_rank=True).to(dev)
    g = torch.Generator().manual_seed(0)
    outputs, ref_v, ref_t, ref_loss, weight = [], [], [], 0.0, 0
    for b in range(3):  # 3 batches of 8 videos x 2 frames (eval_batch_size would be 32)
        video = torch.randn(8, 2, 3, 224, 224, generator=g)
        ids = oracle.tokenize_synthetic(8, (5, 40), seed=100 + b)
        batch = {"video": video.to(dev), "text": {"input_ids": ids.to(dev)}, "video_id": [f"v{b}_{i}" for i in range(8)]}
        out = module.validation_step_end(module.validation_step(batch, b))
        outputs.append(out)
        with torch.inference_mode():
            v, t = ref(video, {"input_ids": ids})
        ref_v.append(v)
        ref_t.append(t)
        ref_loss += float(oracle.ref_nce_loss(ref_batch_scores(v, t, 1 / 0.015))) * 8
        weight += 8
    result = module.validation_epoch_end(outputs)
    assert set(result) == {"r1", "r5", "r10", "mr", "rank", "loss/val"}
    # ranks = the reference ranking of the kernel's own similarity matrix
    ev, et = torch.cat([o[0] for o in outputs]), torch.cat([o[1] for o in outputs])
    own = ops.Similarity(et, ev, 3).scores().cpu()
    expect = oracle.ref_retrieval_metrics(own)
    assert torch.equal(result["rank"].cpu(), expect["rank"])
    assert int(result["mr"]) == int(expect["mr"])
    for k in ("r1", "r5", "r10"):
        assert float(result[k]) == float(expect[k])
    # loss/val: batch-size weighted mean of per-batch NCE at temperature 0.015; bf16 embeddings shift logits by ~1e-1
    assert abs(result["loss/val"] - ref_loss / weight) <= 0.05 * abs(ref_loss / weight)
    # predict_step keys (aligner/video_text_module.py:85-91)
    pred = module.predict_step({"video": batch["video"], "text": batch["text"], "video_id": ["a"] * 8})
    assert set(pred) == {"encoded_videos", "encoded_texts", "video_ids"}


def test_classification_module(pair, dev):
    import oracle
    from oracle.encoder_ref import ref_class_embeddings
    from fitclip_b200 import VideoTextClassificationModule, ops
    ref, enc = pair
    labels, templates = [f"class{i}" for i in range(11)], ["a video of {}", "{} in action", "doing {}"]
    ids = oracle.tokenize_synthetic(11 * 3, (4, 30), seed=7)
    module = VideoTextClassificationModule(enc, labels, templates, tokenized_labels={"input_ids": ids})
    module.on_validation_start()
    with torch.inference_mode():
        ref_labels = ref_class_embeddings(ref, ids, template_count=3)
    assert module.encoded_labels.shape == (11, 512)
    assert torch.nn.functional.cosine_similarity(module.encoded_labels.cpu(), ref_labels).min().item() >= 0.999
    g = torch.Generator().manual_seed(1)
    all_ranks, all_scores, all_y = [], [], []
    for b in range(2):
        video = torch.randn(6, 2, 3, 224, 224, generator=g).to(dev)
        y = torch.randint(0, 11, (6,), generator=g).to(dev)
        logged = module.validation_step({"video": video, "target": (None, y)})
        assert set(logged) == {"a1", "a5", "mr"}
        scores = module(video)
        assert scores.shape == (6, 11)
        all_scores.append(scores.cpu())
        all_y.append(y.cpu())
    result = module.validation_epoch_end()
    s, y = torch.cat(all_scores), torch.cat(all_y)
    assert float(result["a1"]) == float(oracle.ref_accuracy_at_k(s, y, 1))
    assert float(result["a5"]) == float(oracle.ref_accuracy_at_k(s, y, 5))
    assert int(result["mr"]) == int(oracle.ref_median_rank(oracle.ref_rank(s, y)))
    pred = module.predict_step({"video": video, "target": (None, y), "video_id": list("abcdef")})
    assert torch.equal(pred["predictions"].cpu(), all_scores[-1].argmax(dim=-1))


def test_teacher_student_scoring(pair, dev):
    import oracle
    from oracle.encoder_ref import ref_batch_scores
    from fitclip_b200 import B200ClipVideoTextEncoder, TeacherStudentScoringModule
    ref_student, student = pair
    tm = oracle.clip_vit_b_16(seed=1, **TINY)
    ref_teacher = oracle.RefClipVideoTextEncoder(copy.deepcopy(tm))
    teacher = B200ClipVideoTextEncoder(tm.state_dict(), num_frames=2).to(dev)
    module = TeacherStudentScoringModule(student, teacher, init_temperature=0.05).to(dev)
    g = torch.Generator().manual_seed(2)
    video = torch.randn(8, 2, 3, 224, 224, generator=g)
    ids = oracle.tokenize_synthetic(8, (5, 40), seed=9)
    batch = {"video_student": video.to(dev), "text_student": {"input_ids": ids.to(dev)},
             "video_teacher": video.to(dev), "text_teacher": {"input_ids": ids.to(dev)}}
    out = module._step(batch)
    with torch.inference_mode():
        sv, st = ref_student(video, {"input_ids": ids})
        tv, tt = ref_teacher(video, {"input_ids": ids})
    scale = 1 / 0.05
    expect_labeled = float(oracle.ref_nce_loss(ref_batch_scores(sv, st, scale)))
    expect_unlabeled = float(oracle.ref_teacher_student_nce_loss(
        ref_batch_scores(sv, st, scale), ref_batch_scores(tv, tt, scale), reduction="batchmean")) * scale ** 2
    got_labeled = float(module._dataset_step_end(out, "val", "labeled"))
    got_unlabeled = float(module._dataset_step_end(out, "val", "unlabeled"))
    assert abs(got_labeled - expect_labeled) <= 0.02 * abs(expect_labeled)
    assert abs(got_unlabeled - expect_unlabeled) <= 0.1 * abs(expect_unlabeled) + 1e-3
