"""Golden vectors produced by the REFERENCE'S OWN in-tree code (run in the build container, where /root/reference is
mounted: ``python tests/golden/make_reference_golden.py``).  Output: ``tests/golden/reference_outputs.pt``.

The reference cannot be imported as it stands: its third-party dependencies (``overrides``, ``clip``, ``cached_path``,
``torchmetrics``, ``pytorch_lightning``, ``ftfy``, ``timm``) are not installed and there is no network.  This script
puts *stub modules* for exactly those packages into ``sys.modules`` -- no reference file is copied or modified -- and
then imports and RUNS the reference's own modules:

* ``aligner/wise.py``                       ``wise_state_dict`` / ``wise``                       (pure torch)
* ``aligner/loss.py``                       ``nce_loss`` / ``teacher_student_nce_loss`` and the ``_Loss`` modules
* ``aligner/metrics.py``                    ``Rank`` / ``MedianRank`` ``update`` + ``compute``   (stub ``Metric`` base:
                                            ``add_state`` and ``__call__`` only -- the arithmetic is the reference's)
* ``aligner/encoder/clip_video_text_encoder.py``  ``ClipVideoTextEncoder.encode_video / encode_text / forward``, i.e. the
                                            flatten -> encode -> per-frame L2-norm -> frame-mean wrapper, around the
                                            oracle's CLIP module standing in for the un-vendored ``clip.model.CLIP``
* ``aligner/encoder/slip.py``               the in-tree twin of CLIP's text tower (``LayerNorm``, ``QuickGELU``,
                                            ``ResidualAttentionBlock``, ``Transformer``, ``CLIP.encode_text`` with its
                                            causal mask and EOT pooling), loaded with the oracle's text weights
* ``aligner/data/frame_sampler.py``         ``UniformFrameSampler`` (the eval sampler the encoder returns)
* ``aligner/video_text_module.py``, ``aligner/text_video_retrieval.py``   ``validation_step`` -> ``validation_step_end``
                                            (scaled per-batch scores + NCE ``loss/val``) -> ``validation_epoch_end``
                                            (``scores = T @ V.T``, ``target = arange``, ``MedianRank`` / ``Rank``)
* ``aligner/video_text_classification.py``  prompt construction, ``_on_start`` (template mean), ``forward``,
                                            ``validation_step`` (``MedianRank``), ``predict_step``
* ``aligner/teacher_student.py``            ``_step`` + ``_dataset_step_end`` for the labelled and unlabelled split
  (these three run on a stub ``pl.LightningModule`` = ``nn.Module`` + ``log`` / ``all_gather`` / ``print``; ``Recall`` and
  ``Accuracy`` are third-party: the stub restates micro top-k and their values are stored under ``*_restated`` keys)

What stays pinned only against independent implementations (not against reference code, because the code is not in
the reference tree): CLIP's vision tower (``transformers.CLIPModel``, tests/test_oracle_clip.py), torchmetrics'
``Recall`` / ``Accuracy`` definitions, torchvision's transforms (tests/test_oracle_preprocess.py).

tests/test_reference_golden.py checks the oracle (CPU) and the CUDA path (GPU) against the stored outputs."""
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = os.environ.get("FITCLIP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

TINY = dict(embed_dim=64, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=16,
            context_length=16, vocab_size=512, transformer_width=64, transformer_heads=1, transformer_layers=2)


def install_stubs() -> None:
    """Minimal stand-ins for the missing third-party packages -- interface only, no arithmetic of the path."""

    def module(name: str, **attrs) -> types.ModuleType:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def overrides(method=None, *, check_signature=True, check_at_runtime=False):  # decorator with / without arguments
        return method if method is not None else (lambda f: f)

    module("overrides", overrides=overrides)
    module("cached_path", cached_path=lambda p, *a, **k: p)
    module("ftfy", fix_text=lambda t: t)

    class Metric(nn.Module):  # torchmetrics.Metric: the two members aligner/metrics.py uses
        def __init__(self, **kwargs) -> None:
            super().__init__()

        def add_state(self, name, default, dist_reduce_fx=None) -> None:
            setattr(self, name, [] if isinstance(default, list) else default)

        def forward(self, *args, **kwargs):
            self.update(*args, **kwargs)
            return self.compute()

        def clone(self):
            import copy
            return copy.deepcopy(self)

    class TopK(Metric):  # [3P] torchmetrics Recall / Accuracy (multiclass, micro): restated, NOT reference code
        def __init__(self, top_k=None, **kwargs) -> None:
            super().__init__()
            self.top_k = top_k or 1
            self.correct, self.total = 0, 0

        def update(self, preds, target) -> None:
            hits = (preds.topk(self.top_k, dim=1).indices == target.unsqueeze(-1)).any(dim=1)
            self.correct += int(hits.sum())
            self.total += int(hits.numel())

        def compute(self):
            return torch.tensor(self.correct / self.total, dtype=torch.float32)

    def unavailable(*_a, **_k):
        raise RuntimeError("third-party class not available in the stub")

    module("torchmetrics", Metric=Metric, Recall=TopK, Accuracy=TopK)

    registry = module("timm.models.registry", register_model=lambda f: f)
    models = module("timm.models", registry=registry, vision_transformer=types.SimpleNamespace())
    module("timm", models=models, create_model=unavailable)

    clip_model = module("clip.model", CLIP=oracle.CLIP)
    clip_clip = module("clip.clip", load=unavailable, tokenize=lambda texts, truncate=True: synthetic_tokenize(texts),
                       _tokenizer=None, model=clip_model)
    pkg = module("clip", clip=clip_clip, model=clip_model)
    pkg.__path__ = []  # a package, so that `from clip import clip` / `from clip.model import CLIP` resolve

    def apply_to_collection(data, dtype, function, *args, **kwargs):  # [3P] PL utility, restated
        if isinstance(data, dtype):
            return function(data, *args, **kwargs)
        if isinstance(data, dict) or hasattr(data, "items"):
            return {k: apply_to_collection(v, dtype, function, *args, **kwargs) for k, v in data.items()}
        if isinstance(data, (list, tuple)):
            return type(data)(apply_to_collection(v, dtype, function, *args, **kwargs) for v in data)
        return data

    class LightningModule(nn.Module):  # [3P] the members the reference modules touch during evaluation
        def __init__(self) -> None:
            super().__init__()
            self.logged = []
            self.global_step = 0
            self._current_dataloader_idx = None

        def log(self, name, value, **kwargs) -> None:
            value = value.compute() if hasattr(value, "compute") else value
            self.logged.append((name, value.detach().clone() if isinstance(value, torch.Tensor) else value,
                                kwargs.get("batch_size")))

        def all_gather(self, data, group=None, sync_grads=False):  # one device: identity
            return data

        def print(self, *args, **kwargs) -> None:
            pass

    class RichProgressBar:
        class _Progress:
            def add_task(self, **kwargs):
                return 0

            def update(self, *args, **kwargs) -> None:
                pass

        def __init__(self) -> None:
            self.progress = self._Progress()

    apply_func = module("pytorch_lightning.utilities.apply_func", apply_to_collection=apply_to_collection)
    utilities = module("pytorch_lightning.utilities", apply_func=apply_func)
    callbacks = module("pytorch_lightning.callbacks", RichProgressBar=RichProgressBar)
    pl = module("pytorch_lightning", LightningModule=LightningModule, utilities=utilities, callbacks=callbacks)
    pl.__path__ = []


def synthetic_tokenize(texts, context_length: int = 16, vocab_size: int = 512, truncate: bool = True) -> torch.Tensor:
    """Stand-in for ``clip.tokenize`` (third-party, vocabulary file not on disk): a deterministic word hash with CLIP's
    framing -- SOT = vocab-2, EOT = vocab-1 (the row maximum, which is what the EOT pooling looks for), zero padding,
    truncation keeps EOT last."""
    import zlib
    rows = []
    for text in texts:
        ids = [vocab_size - 2] + [1 + zlib.crc32(w.encode()) % (vocab_size - 3) for w in text.lower().split()] + \
              [vocab_size - 1]
        if len(ids) > context_length:
            ids = ids[:context_length]
            ids[-1] = vocab_size - 1
        rows.append(ids + [0] * (context_length - len(ids)))
    return torch.tensor(rows, dtype=torch.int32)


def main() -> None:
    assert os.path.isdir(REFERENCE), f"{REFERENCE} is not mounted: this script only runs in the build container"
    torch.set_num_threads(1)
    install_stubs()
    sys.path.insert(0, REFERENCE)
    from aligner import loss as ref_loss  # noqa: E402  (reference modules)
    from aligner import metrics as ref_metrics  # noqa: E402
    from aligner import wise as ref_wise  # noqa: E402
    from aligner.data.frame_sampler import UniformFrameSampler  # noqa: E402
    from aligner.encoder import slip as ref_slip  # noqa: E402
    from aligner.encoder.clip_video_text_encoder import ClipVideoTextEncoder  # noqa: E402

    out = {"config": TINY, "reference_files": [
        "aligner/wise.py", "aligner/loss.py", "aligner/metrics.py", "aligner/encoder/clip_video_text_encoder.py",
        "aligner/encoder/slip.py", "aligner/data/frame_sampler.py"]}
    g = torch.Generator().manual_seed(20221118)

    # ---- the wrapper (clip_video_text_encoder.py:68-94) around the oracle's CLIP
    m1 = oracle.clip_vit_b_16(seed=0, **TINY)
    m2 = oracle.clip_vit_b_16(seed=1, **TINY)
    out["state_dict_1"] = {k: v.clone() for k, v in m1.state_dict().items()}
    out["state_dict_2"] = {k: v.clone() for k, v in m2.state_dict().items()}
    video = torch.randn(12, 3, 3, 32, 32, generator=g)
    ids = oracle.tokenize_synthetic(12, (3, 16), seed=77, context_length=16, vocab_size=512)
    enc1 = ClipVideoTextEncoder(m1, num_frames=3)
    enc2 = ClipVideoTextEncoder(m2, num_frames=3)
    assert not hasattr(enc1.model, "logit_scale")  # :75-77
    with torch.inference_mode():
        v, t = enc1(video=video, text={"input_ids": ids})
    out.update(video=video, input_ids=ids, wrapper_video_emb=v.clone(), wrapper_text_emb=t.clone(),
               wrapper_param_names=[n for n, _ in enc1.named_parameters()])

    # ---- the in-tree twin of the text tower (slip.py:350-480) with the same text weights
    twin = ref_slip.CLIP(embed_dim=64, vision_width=64, vision_model=nn.Identity(), context_length=16, vocab_size=512,
                         transformer_width=64, transformer_heads=1, transformer_layers=2)
    text_sd = {k: v for k, v in out["state_dict_1"].items() if not k.startswith("visual.") and k != "logit_scale"}
    missing, unexpected = twin.load_state_dict(text_sd, strict=False)
    assert not unexpected and set(missing) <= {"image_projection", "logit_scale"}, (missing, unexpected)
    with torch.inference_mode():
        out["slip_text_features"] = twin.eval().encode_text(ids.long()).clone()  # un-normalised, EOT-pooled + projected

    # ---- WiSE (wise.py:10-23) on the two encoders
    keep = ("model.text_projection", "model.visual.conv1.weight", "model.visual.transformer.resblocks.1.ln_2.bias",
            "model.transformer.resblocks.0.attn.in_proj_weight", "model.token_embedding.weight")
    for w in (0.4, 0.5):
        wsd = ref_wise.wise_state_dict(enc1, enc2, weight_for_2=w)
        assert set(wsd) == {n for n, _ in enc1.named_parameters()}
        out[f"wise_{w}_state_dict"] = {k: wsd[k].detach().clone() for k in keep}  # a sample keeps the fixture small
    wenc = ref_wise.wise(enc1, enc2, weight_for_2=0.4)
    with torch.inference_mode():
        wv, wt = wenc(video=video, text={"input_ids": ids})
    out.update(wise_video_emb=wv.clone(), wise_text_emb=wt.clone())

    # ---- losses (loss.py:13-65)
    scores = torch.randn(9, 9, generator=g) * 4
    teacher = torch.randn(9, 9, generator=g) * 4
    out["loss_scores"], out["loss_teacher_scores"] = scores, teacher
    for red in ("mean", "sum", "none"):
        out[f"nce_{red}"] = ref_loss.nce_loss(scores, reduction=red).clone()
    for red in ("mean", "sum", "batchmean"):
        out[f"ts_nce_{red}"] = ref_loss.teacher_student_nce_loss(scores, teacher, reduction=red).clone()
    out["nce_module"] = ref_loss.NCELoss()(scores).clone()
    out["ts_nce_module"] = ref_loss.TeacherStudentNCELoss(reduction="batchmean")(scores, teacher).clone()
    out["similarity_loss"] = ref_loss.SimilarityLoss()(scores).clone()

    # ---- rank metrics (metrics.py:6-36) on tie-free score matrices, fed in two batches like a validation epoch
    rank_cases = []
    for rows, cols in ((40, 40), (17, 101), (64, 9)):
        s = torch.randn(rows, cols, generator=g)
        tgt = torch.arange(rows) % cols if rows != cols else torch.arange(rows)
        rank_m, med_m = ref_metrics.Rank(), ref_metrics.MedianRank()
        half = rows // 2
        for a, b in ((0, half), (half, rows)):
            rank_m.update(s[a:b], tgt[a:b])
            med_m.update(s[a:b], tgt[a:b])
        rank_cases.append({"scores": s, "target": tgt, "ranks": rank_m.compute().clone(),
                           "median_rank": med_m.compute().clone()})
    out["rank_cases"] = rank_cases

    # ---- the eval frame sampler (frame_sampler.py:32-41)
    sampler_cases = []
    for max_frames, (a, b) in ((4, (0, 99)), (8, (10, 300)), (4, (5, 6)), (4, (0, 0)), (3, (7, 1000))):
        idx = [int(i) for i in UniformFrameSampler(max_frames)(a, b, 30.0)]
        sampler_cases.append({"max_frames": max_frames, "start": a, "end": b, "indices": idx})
    out["uniform_sampler_cases"] = sampler_cases

    # ---- retrieval evaluation flow (video_text_module.py:38-44, text_video_retrieval.py:40-98) in batches of 6
    from aligner.teacher_student import TeacherStudentLightningModule  # noqa: E402
    from aligner.text_video_retrieval import TextVideoRetrievalLightningModule  # noqa: E402
    from aligner.video_text_classification import VideoTextClassificationLightningModule  # noqa: E402
    module = TextVideoRetrievalLightningModule(encoder=enc1, init_temperature=0.015, fit_temperature=False,
                                               compute_rank=True)
    outputs = []
    with torch.inference_mode():
        for lo in (0, 6):
            batch = {"video": video[lo:lo + 6], "text": {"input_ids": ids[lo:lo + 6]},
                     "video_id": [f"v{i}" for i in range(lo, lo + 6)]}
            outputs.append(module.validation_step_end(module.validation_step(batch)))
        module.validation_epoch_end(outputs)
    logged = module.logged
    out["retrieval"] = {
        "init_temperature": 0.015, "batch_losses": [v for n, v, _ in logged if n == "loss/val"],
        "batch_sizes": [b for n, _, b in logged if n == "loss/val"],
        "mr": next(v for n, v, _ in logged if n == "mr"), "rank": next(v for n, v, _ in logged if n == "rank"),
        "r1_restated": next(v for n, v, _ in logged if n == "r1"), "r5_restated": next(v for n, v, _ in logged if n == "r5"),
        "encoded_videos": torch.cat([o[0] for o in outputs]).clone(), "encoded_texts": torch.cat([o[1] for o in outputs]).clone()}

    # ---- the same flow with TWO datasets (text_video_retrieval.py:28-37, 60-65, 84-93): per-dataset metric clones,
    # `dataloader_idx`, `loss/val_{name}`; dataset "a" = all 12 samples, dataset "b" = samples 1..11 (Recall(top_k=10) needs more than 10
    # candidates per dataset, as torchmetrics does), one batch each
    multi = TextVideoRetrievalLightningModule(encoder=enc1, init_temperature=0.015, fit_temperature=False,
                                              dataset_names=["a", "b"])
    multi_outputs = []
    with torch.inference_mode():
        for idx, (lo, hi) in enumerate(((0, 12), (1, 12))):
            batch = {"video": video[lo:hi], "text": {"input_ids": ids[lo:hi]}}
            multi_outputs.append([multi.validation_step_end(multi.validation_step(batch, 0, idx))])
        multi.validation_epoch_end(multi_outputs)
    out["retrieval_multi"] = {
        "splits": [(0, 12), (1, 12)], "dataset_names": ["a", "b"],
        "logged": {n: (v.clone() if isinstance(v, torch.Tensor) else v) for n, v, _ in multi.logged
                   if n.startswith("loss/val") or n.startswith("mr_")},
        "logged_names": sorted({n for n, _, _ in multi.logged})}

    # ---- zero-shot classification flow (video_text_classification.py:30-140)
    labels = ["archery", "baby crawling", "cutting in kitchen", "drumming", "fencing", "golf swing", "knitting"]
    templates = ["a video of a person {}.", "{} in action", "someone is {} here"]
    cls = VideoTextClassificationLightningModule(enc1, labels=labels, templates=templates, init_temperature=0.015,
                                                 fit_temperature=False)
    cls.trainer = types.SimpleNamespace(callbacks=[sys.modules["pytorch_lightning.callbacks"].RichProgressBar()],
                                        is_global_zero=True)
    label_ids = torch.tensor([3, 0, 6, 6, 2, 5, 1, 4, 0, 3, 5, 2])
    with torch.inference_mode():
        cls.on_validation_start()
        cls_scores = cls(video)
        cls.validation_step({"video": video, "target": (["x"] * 12, label_ids)})
        pred = cls.predict_step({"video": video, "target": (["x"] * 12, label_ids), "video_id": list(range(12))})
    out["classification"] = {
        "labels": labels, "templates": templates, "tokenized_prompts": cls.tokenized_labels["input_ids"].detach().clone(),
        "encoded_labels": cls.encoded_labels.detach().clone(), "scores": cls_scores.clone(), "label_ids": label_ids,
        "mr": next(v for n, v, _ in cls.logged if n == "mr"), "a1_restated": next(v for n, v, _ in cls.logged if n == "a1"),
        "a5_restated": next(v for n, v, _ in cls.logged if n == "a5"), "predictions": pred["predictions"].clone()}

    # ---- teacher-student scoring (teacher_student.py:93-96,142-173): labelled -> NCE, unlabelled -> KL * scale^2
    ts = TeacherStudentLightningModule(encoder=enc1, teacher=enc2, init_temperature=0.015, fit_temperature=False)
    with torch.inference_mode():
        step = ts._step({"video_student": video, "text_student": {"input_ids": ids}, "video_teacher": video,
                         "text_teacher": {"input_ids": ids}})
        ts._dataset_step_end(step, split="val", dataset_name="labeled")
        ts._dataset_step_end(step, split="val", dataset_name="unlabeled")
    out["teacher_student"] = {"init_temperature": 0.015,
                              "loss_labeled": next(v for n, v, _ in ts.logged if n == "loss/val_labeled"),
                              "loss_unlabeled": next(v for n, v, _ in ts.logged if n == "loss/val_unlabeled"),
                              "teacher_video_emb": step[1][0].clone(), "teacher_text_emb": step[1][1].clone()}
    out["reference_files"] += ["aligner/video_text_module.py", "aligner/text_video_retrieval.py",
                               "aligner/video_text_classification.py", "aligner/teacher_student.py",
                               "util/tensor_utils.py"]

    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB); torch {torch.__version__}")


if __name__ == "__main__":
    main()
