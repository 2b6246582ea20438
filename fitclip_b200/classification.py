"""Zero-shot video classification on the native encoder: the hooks of ``VideoTextClassificationLightningModule``
(``aligner/video_text_classification.py:30-140``) without Lightning.

``labels x templates`` prompts are encoded once (``_on_start``, ``:69-96``), averaged over templates WITHOUT
re-normalising (``:95``); ``forward(video) = encode_video(video) @ encoded_labels.T`` (``:115-116``); metrics ``a1``,
``a5`` (micro top-k accuracy) and ``mr`` (median rank) are updated per batch (``:119-126``).  Rows are independent, so
multi-GPU evaluation shards videos and only concatenates the integer ranks (``dist_reduce_fx="cat"``)."""
from __future__ import annotations

from typing import Any, Dict, Iterable, Mapping, Optional

import torch
from torch import nn

from . import ops
from .api import VideoTextEncoder
from .metrics import Accuracy, MedianRank, Rank


class VideoTextClassificationModule(nn.Module):
    def __init__(self, encoder: VideoTextEncoder, labels: Iterable[str], templates: Optional[Iterable[str]] = None,
                 return_metrics_by_class: bool = False, tokenized_labels: Optional[Mapping[str, torch.Tensor]] = None,
                 group=None, similarity_terms: int = 3) -> None:
        """``tokenized_labels`` bypasses the tokenizer (synthetic runs: the CLIP BPE vocabulary is not on this image);
        it must hold ``len(labels) * len(templates)`` rows ordered label-major like the reference (``:47``)."""
        super().__init__()
        self.encoder = encoder
        labels = list(labels)
        self.label_count = len(labels)
        if templates:
            templates = list(templates)
            self.template_count = len(templates)
            prompts = [template.format(label) for label in labels for template in templates]
        else:
            self.template_count = 1
            prompts = labels
        if tokenized_labels is None:
            tokenized_labels = encoder.get_tokenizer()(prompts)
        n = next(iter(tokenized_labels.values())).shape[0]
        assert n == self.label_count * self.template_count, (n, self.label_count, self.template_count)
        self.tokenized_labels = {k: v for k, v in tokenized_labels.items()}
        self.encoded_labels: Optional[torch.Tensor] = None
        self.similarity_terms = similarity_terms
        self.metrics: Dict[str, Rank] = {"a1": Accuracy(process_group=group), "a5": Accuracy(top_k=5, process_group=group),
                                         "mr": MedianRank(process_group=group)}
        # per-class accuracy (:62-66): (rank, label) pairs are kept in two "cat" states and reduced ONCE per epoch, so
        # every rank issues the same collectives whatever classes its shard happened to contain
        self.metrics_by_class = ({"rank": Rank(process_group=group), "label": Rank(process_group=group)}
                                 if return_metrics_by_class else None)

    def _on_start(self) -> None:
        device = next(self.encoder.parameters()).device
        tokens = {k: v.to(device) for k, v in self.tokenized_labels.items()}
        encoded = self.encoder.encode_text(tokens)  # per-caption results do not depend on the batch split (:83-84)
        self.encoded_labels = encoded.reshape(-1, self.template_count, encoded.shape[1]).mean(dim=1).contiguous()

    on_validation_start = on_test_start = on_predict_start = _on_start

    def forward(self, video: torch.Tensor) -> torch.Tensor:
        if self.encoded_labels is None:
            self._on_start()
        encoded_video = self.encoder.encode_video(video)
        return ops.Similarity(encoded_video, self.encoded_labels, self.similarity_terms).scores()

    def validation_step(self, batch: Mapping[str, Any], _batch_idx: int = 0) -> Dict[str, torch.Tensor]:
        scores = self(batch["video"])
        label_id = batch["target"][1]
        logged = {}
        ranks = ops.rank_from_scores(scores, label_id)  # one pass over the scores serves a1, a5 and mr
        for name, metric in self.metrics.items():
            metric.update_from_ranks(ranks, scores.shape[1])
            logged[name] = metric._compute_from(ranks)
        if self.metrics_by_class is not None:
            self.metrics_by_class["rank"].update_from_ranks(ranks, scores.shape[1])
            self.metrics_by_class["label"].update_from_ranks(label_id.to(ranks.device), self.label_count)
        return logged

    def validation_epoch_end(self, _outputs=None) -> Dict[str, torch.Tensor]:
        result = {name: metric.compute() for name, metric in self.metrics.items()}
        if self.metrics_by_class is not None:
            by_class = self.metrics_by_class
            if not by_class["rank"].ranks:  # a rank whose shard was empty still takes part in the two gathers
                device = next(self.encoder.parameters()).device
                for m in by_class.values():
                    m.update_from_ranks(torch.empty(0, dtype=torch.int64, device=device), self.label_count)
            ranks, labels = by_class["rank"].compute(), by_class["label"].compute()
            hits = torch.bincount(labels[ranks < 1], minlength=self.label_count).to(torch.float32)
            seen = torch.bincount(labels, minlength=self.label_count)
            for k in torch.nonzero(seen).flatten().tolist():  # classes without samples log nothing, as in the reference
                result[f"a1_{k}"] = hits[k] / seen[k]
            for m in by_class.values():
                m.reset()
        for metric in self.metrics.values():
            metric.reset()
        return result

    def predict_step(self, batch: Mapping[str, Any], _batch_idx: int = 0) -> Mapping[str, Any]:
        values, indices = ops.topk_rows(self(batch["video"]), 1)  # argmax, lowest index on ties (:135-140)
        return {"predictions": indices[:, 0].long(), "labels": batch["target"][1], "video_ids": batch.get("video_id")}
