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


# ------------------------------------------------------------------------- module flows run by the reference's own classes
def test_oracle_retrieval_flow_matches_reference_module(ref):
    """TextVideoRetrievalLightningModule (validation_step -> validation_step_end -> validation_epoch_end) as executed by
    the reference: per-batch scaled NCE loss, then ranks of `T @ V.T` against `arange`."""
    from oracle.encoder_ref import ref_batch_scores, ref_retrieval_scores
    r = ref["retrieval"]
    enc = _oracle_encoder(ref)
    scale = 1.0 / r["init_temperature"]
    vs, ts = [], []
    with torch.inference_mode():
        for i, lo in enumerate((0, 6)):
            v, t = enc(ref["video"][lo:lo + 6], {"input_ids": ref["input_ids"][lo:lo + 6]})
            loss = oracle.ref_nce_loss(ref_batch_scores(v, t, scale))
            assert abs(loss.item() - r["batch_losses"][i].item()) <= 1e-4 * abs(r["batch_losses"][i].item())
            vs.append(v)
            ts.append(t)
    scores = ref_retrieval_scores(torch.cat(ts), torch.cat(vs))
    ranks = oracle.ref_rank(scores, torch.arange(12))
    assert torch.equal(ranks, r["rank"])
    assert int(oracle.ref_median_rank(ranks)) == int(r["mr"])
    m = oracle.ref_retrieval_metrics(scores)
    assert abs(float(m["r1"]) - float(r["r1_restated"])) < 1e-7 and abs(float(m["r5"]) - float(r["r5_restated"])) < 1e-7


def test_oracle_classification_flow_matches_reference_module(ref):
    from oracle.encoder_ref import ref_class_embeddings
    c = ref["classification"]
    enc = _oracle_encoder(ref)
    with torch.inference_mode():
        emb = ref_class_embeddings(enc, c["tokenized_prompts"], len(c["templates"]))
        scores = enc.encode_video(ref["video"]) @ emb.T
    assert torch.allclose(emb, c["encoded_labels"], atol=1e-6, rtol=0)
    assert torch.allclose(scores, c["scores"], atol=1e-6, rtol=0)
    ranks = oracle.ref_rank(c["scores"], c["label_ids"])
    assert int(oracle.ref_median_rank(ranks)) == int(c["mr"])
    assert torch.equal(c["scores"].argmax(dim=-1), c["predictions"])
    assert abs(float(oracle.ref_accuracy_at_k(c["scores"], c["label_ids"], 1)) - float(c["a1_restated"])) < 1e-7


def test_oracle_teacher_student_scoring_matches_reference_module(ref):
    from oracle.encoder_ref import ref_batch_scores
    t = ref["teacher_student"]
    scale = 1.0 / t["init_temperature"]
    student, teacher = _oracle_encoder(ref, "state_dict_1"), _oracle_encoder(ref, "state_dict_2")
    with torch.inference_mode():
        sv, st = student(ref["video"], {"input_ids": ref["input_ids"]})
        tv, tt = teacher(ref["video"], {"input_ids": ref["input_ids"]})
    assert torch.allclose(tv, t["teacher_video_emb"], atol=1e-6) and torch.allclose(tt, t["teacher_text_emb"], atol=1e-6)
    scores, teacher_scores = ref_batch_scores(sv, st, scale), ref_batch_scores(tv, tt, scale)
    labeled = oracle.ref_nce_loss(scores)
    unlabeled = oracle.ref_teacher_student_nce_loss(scores, teacher_scores, reduction="batchmean") * scale ** 2
    assert abs(labeled.item() - t["loss_labeled"].item()) <= 1e-4 * abs(t["loss_labeled"].item())
    assert abs(unlabeled.item() - t["loss_unlabeled"].item()) <= 1e-3 * abs(t["loss_unlabeled"].item())


@pytest.mark.gpu
def test_cuda_retrieval_module_matches_reference_module(ref, dev):
    from fitclip_b200 import B200ClipVideoTextEncoder, TextVideoRetrievalModule
    r = ref["retrieval"]
    enc = B200ClipVideoTextEncoder(ref["state_dict_1"], num_frames=3).to(dev)
    module = TextVideoRetrievalModule(enc, init_temperature=r["init_temperature"], fit_temperature=False,
                                      compute_rank=True).to(dev)
    outputs = []
    with torch.inference_mode():
        for lo in (0, 6):
            batch = {"video": ref["video"][lo:lo + 6].to(dev), "text": {"input_ids": ref["input_ids"][lo:lo + 6].to(dev)},
                     "video_id": [f"v{i}" for i in range(lo, lo + 6)]}
            outputs.append(module.validation_step_end(module.validation_step(batch)))
        result = module.validation_epoch_end(outputs)
    # loss/val: PL's batch-size weighted mean of the per-batch losses; bf16 encoder vs fp32 reference at scale 66.7
    expect = sum(float(l) * b for l, b in zip(r["batch_losses"], r["batch_sizes"])) / sum(r["batch_sizes"])
    assert abs(float(result["loss/val"]) - expect) <= 3e-2 * abs(expect)
    # the epoch-end logic on the REFERENCE'S embeddings: exact ranks / MdR
    injected = [(r["encoded_videos"][lo:lo + 6].to(dev), r["encoded_texts"][lo:lo + 6].to(dev)) for lo in (0, 6)]
    exact = module._validate_dataset(injected)
    assert torch.equal(exact["rank"].cpu(), r["rank"])
    assert int(exact["mr"]) == int(r["mr"])
    assert abs(float(exact["r1"]) - float(r["r1_restated"])) < 1e-7 and abs(float(exact["r5"]) - float(r["r5_restated"])) < 1e-7


def _multi_dataset_flow(module, encode_inputs, ref):
    """Drives a TextVideoRetrievalModule the way the golden script drove the reference's module in multi-dataset mode."""
    m = ref["retrieval_multi"]
    outputs = []
    for idx, (lo, hi) in enumerate(m["splits"]):
        batch = encode_inputs(lo, hi)
        outputs.append([module.validation_step_end(module.validation_step(batch, 0, idx))])
    return module.validation_epoch_end(outputs)


def test_multi_dataset_module_matches_reference_module_on_cpu(ref):
    """Multi-dataset validation (text_video_retrieval.py:28-37, 60-65, 84-93) against what the REFERENCE'S module logged:
    our module's host logic on the reference's own (oracle) embeddings with CPU stand-ins for the kernels -- metric names,
    per-dataset MdR exactly, per-dataset loss/val to fp32 precision."""
    from fitclip_b200 import TextVideoRetrievalModule, ops
    from test_distributed_gloo import CpuSimilarity, _StubEncoder, _cpu_metrics_from_ranks
    m, r = ref["retrieval_multi"], ref["retrieval"]
    backup = ops.metrics_from_ranks
    ops.metrics_from_ranks = _cpu_metrics_from_ranks
    try:
        module = TextVideoRetrievalModule(_StubEncoder(), init_temperature=r["init_temperature"], fit_temperature=False,
                                          dataset_names=m["dataset_names"], similarity_factory=CpuSimilarity,
                                          nce_loss_fn=oracle.ref_nce_loss)
        result = _multi_dataset_flow(module, lambda lo, hi: {"video": r["encoded_videos"][lo:hi],
                                                             "text": {"input_ids": r["encoded_texts"][lo:hi]}}, ref)
    finally:
        ops.metrics_from_ranks = backup
    assert sorted(result) == m["logged_names"]
    for name in m["dataset_names"]:
        assert int(result[f"mr_{name}"]) == int(m["logged"][f"mr_{name}"])
        assert abs(float(result[f"loss/val_{name}"]) - float(m["logged"][f"loss/val_{name}"])) <= 1e-4


@pytest.mark.gpu
def test_cuda_multi_dataset_module_matches_reference_module(ref, dev):
    from fitclip_b200 import B200ClipVideoTextEncoder, TextVideoRetrievalModule
    m, r = ref["retrieval_multi"], ref["retrieval"]
    enc = B200ClipVideoTextEncoder(ref["state_dict_1"], num_frames=3).to(dev)
    module = TextVideoRetrievalModule(enc, init_temperature=r["init_temperature"], fit_temperature=False,
                                      dataset_names=m["dataset_names"]).to(dev)
    with torch.inference_mode():
        result = _multi_dataset_flow(module, lambda lo, hi: {"video": ref["video"][lo:hi].to(dev),
                                                             "text": {"input_ids": ref["input_ids"][lo:hi].to(dev)}}, ref)
    assert sorted(result) == m["logged_names"]
    for name in m["dataset_names"]:  # bf16 encoder vs fp32 reference at scale 66.7
        expect = float(m["logged"][f"loss/val_{name}"])
        assert abs(float(result[f"loss/val_{name}"]) - expect) <= 3e-2 * abs(expect)
    # the epoch-end logic on the REFERENCE'S embeddings: exact per-dataset MdR
    module2 = TextVideoRetrievalModule(enc, init_temperature=r["init_temperature"], fit_temperature=False,
                                       dataset_names=m["dataset_names"]).to(dev)
    for name, (lo, hi) in zip(m["dataset_names"], m["splits"]):
        exact = module2._validate_dataset([(r["encoded_videos"][lo:hi].to(dev), r["encoded_texts"][lo:hi].to(dev))],
                                          dataset_name=name)
        assert int(exact[f"mr_{name}"]) == int(m["logged"][f"mr_{name}"])


@pytest.mark.gpu
def test_cuda_classification_module_matches_reference_module(ref, dev):
    from fitclip_b200 import B200ClipVideoTextEncoder, ops
    from fitclip_b200.classification import VideoTextClassificationModule
    c = ref["classification"]
    enc = B200ClipVideoTextEncoder(ref["state_dict_1"], num_frames=3).to(dev)
    module = VideoTextClassificationModule(enc, labels=c["labels"], templates=c["templates"],
                                           tokenized_labels={"input_ids": c["tokenized_prompts"]})
    with torch.inference_mode():
        module.on_validation_start()
        scores = module(ref["video"].to(dev))
    assert F.cosine_similarity(module.encoded_labels.cpu(), c["encoded_labels"]).min().item() >= 0.999
    assert (scores.cpu() - c["scores"]).abs().max().item() <= 2e-2
    # metric / prediction logic on the REFERENCE'S scores: exact
    ref_scores = c["scores"].to(dev)
    ranks = ops.rank_from_scores(ref_scores, c["label_ids"].to(dev))
    assert int(ranks.cpu().median()) + 1 == int(c["mr"])
    _, idx = ops.topk_rows(ref_scores, 1)
    assert torch.equal(idx[:, 0].long().cpu(), c["predictions"])


@pytest.mark.gpu
def test_cuda_teacher_student_module_matches_reference_module(ref, dev):
    from fitclip_b200 import B200ClipVideoTextEncoder
    from fitclip_b200.teacher_student import TeacherStudentScoringModule
    t, r = ref["teacher_student"], ref["retrieval"]
    student = B200ClipVideoTextEncoder(ref["state_dict_1"], num_frames=3).to(dev)
    teacher = B200ClipVideoTextEncoder(ref["state_dict_2"], num_frames=3).to(dev)
    module = TeacherStudentScoringModule(student, teacher, init_temperature=t["init_temperature"]).to(dev)
    # the scoring arithmetic on the REFERENCE'S embeddings
    injected = ((r["encoded_videos"].to(dev), r["encoded_texts"].to(dev)),
                (t["teacher_video_emb"].to(dev), t["teacher_text_emb"].to(dev)))
    labeled = float(module._dataset_step_end(injected, dataset_name="labeled"))
    unlabeled = float(module._dataset_step_end(injected, dataset_name="unlabeled"))
    assert abs(labeled - float(t["loss_labeled"])) <= 1e-3 * abs(float(t["loss_labeled"]))
    assert abs(unlabeled - float(t["loss_unlabeled"])) <= 1e-3 * abs(float(t["loss_unlabeled"]))
    # and end to end through both native encoders (bf16 embeddings, scores scaled by 66.7, KL scaled by 66.7^2)
    batch = {"video_student": ref["video"].to(dev), "text_student": {"input_ids": ref["input_ids"].to(dev)},
             "video_teacher": ref["video"].to(dev), "text_teacher": {"input_ids": ref["input_ids"].to(dev)}}
    with torch.inference_mode():
        step = module._step(batch)
        labeled = float(module._dataset_step_end(step, dataset_name="labeled"))
        unlabeled = float(module._dataset_step_end(step, dataset_name="unlabeled"))
    assert abs(labeled - float(t["loss_labeled"])) <= 5e-2 * abs(float(t["loss_labeled"]))
    assert abs(unlabeled - float(t["loss_unlabeled"])) <= 1e-1 * abs(float(t["loss_unlabeled"]))
