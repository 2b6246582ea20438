"""Similarity + rank side of the hot path, single- and multi-GPU.

* :func:`retrieval_ranks` -- the fused path: ``scores = texts @ videos.T`` (``aligner/text_video_retrieval.py:74``),
  ``target = arange`` (``:76``) and ``Rank.update`` (``aligner/metrics.py:16-19``) without ever building the
  ``Nt x Nv`` matrix, sharded by video columns across ranks (one process per GPU).
* :class:`TextVideoRetrievalModule` -- the hooks of ``TextVideoRetrievalLightningModule``
  (``text_video_retrieval.py:16-98``) and ``VideoTextLightningModule`` (``video_text_module.py:25-91``) without
  Lightning: ``forward`` / ``validation_step`` / ``validation_step_end`` / ``validation_epoch_end`` / ``predict_step``
  and the logged keys ``loss/val``, ``r1``, ``r5``, ``r10``, ``mr``.

Sharding: rank r owns a contiguous slice of videos and captions.  The only exchanges are an
all-gather of text embeddings, an all-reduce(sum) of fp32 target scores (non-owners contribute exact zeros, so the sum
is exact) and an all-reduce(sum) of int32 rank counts -- results are identical on every rank and identical to the
single-GPU run, with no padding or duplicated samples (unlike the reference's DistributedSampler,
``config/trainer.yaml:31``)."""
from __future__ import annotations

import math
from typing import Any, Callable, Dict, List, Mapping, MutableMapping, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .api import TYPE_OUTPUT, VideoTextEncoder
from .metrics import MedianRank, Rank, Recall


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`: ceil(n / world) per rank, last shards may be short/empty."""
    per = -(-n // world)
    return min(n, rank * per), min(n, (rank + 1) * per)


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_rows(x: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """Concatenate per-rank row blocks of different heights (pad to the tallest, gather, strip).
    Returns the gathered tensor and the per-rank heights."""
    world, _ = _world(group)
    if world == 1:
        return x, [x.shape[0]]
    n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    padded = x.new_zeros((cap, *x.shape[1:]))
    padded[:x.shape[0]] = x
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)]), sizes


def retrieval_ranks(text_local: torch.Tensor, video_local: torch.Tensor, group=None, terms: int = 3,
                    target_local: Optional[torch.Tensor] = None,
                    similarity_factory: Callable[..., Any] = ops.Similarity) -> torch.Tensor:
    """0-based rank of every query's target video among ALL videos, int64 ``(Nt_total,)``, same on every rank.

    ``text_local (nt_r, E)`` / ``video_local (nv_r, E)``: this rank's fp32 embeddings, contiguous shards in rank
    order.  ``target_local``: global video index per local query (default: query i <-> video i).
    ``similarity_factory`` exists so the host-side sharding logic can be exercised without a GPU (tests only)."""
    world, rank = _world(group)
    text_all, _ = all_gather_rows(text_local.contiguous(), group)
    if world > 1:
        nv = torch.tensor([video_local.shape[0]], device=video_local.device, dtype=torch.int64)
        sizes = [torch.zeros_like(nv) for _ in range(world)]
        dist.all_gather(sizes, nv, group=group)
        col_offset = int(sum(int(s.item()) for s in sizes[:rank]))
    else:
        col_offset = 0
    if target_local is None:
        target = torch.arange(text_all.shape[0], device=text_all.device, dtype=torch.int32)
    else:
        target, _ = all_gather_rows(target_local.to(torch.int32).contiguous(), group)
    nt = text_all.shape[0]
    if video_local.shape[0] > 0:
        sim = similarity_factory(text_all, video_local.contiguous(), terms)
        tscore = sim.target_scores(target, col_offset)
    else:  # an empty shard still has to take part in the collectives
        sim = None
        tscore = torch.zeros(nt, device=text_all.device, dtype=torch.float32)
    if world > 1:
        dist.all_reduce(tscore, op=dist.ReduceOp.SUM, group=group)
    counts = sim.counts(target, tscore, col_offset) if sim is not None else \
        torch.zeros(nt, device=text_all.device, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts.to(torch.int64)


def retrieval_topk(text_local: torch.Tensor, video_local: torch.Tensor, k: int = 10, group=None, terms: int = 3,
                   row_chunk: int = 8192, similarity_factory: Callable[..., Any] = ops.Similarity,
                   topk_fn: Callable[..., Tuple[torch.Tensor, torch.Tensor]] = ops.topk_rows
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The ``k`` best videos of every query over ALL videos: ``(values (Nt_total, k) fp32, indices (Nt_total, k) int64
    global video index)``, identical on every rank -- what ``torch.topk(texts @ videos.T, k)`` returns on one device
    (``Recall(top_k)`` looks at exactly this set, ``aligner/text_video_retrieval.py:21``), with ties broken towards the
    lower video index.

    Sharding as in :func:`retrieval_ranks`: text embeddings are all-gathered, every rank scores its own video-column
    slab in chunks of ``row_chunk`` queries (so the ``Nt x Nv`` matrix never exists), keeps a local top-k per query, and
    the ``world x k`` candidates per query are all-gathered and merged with the same top-k kernel.  ``k`` may exceed a
    shard's (or the gallery's) size: missing slots are ``-inf`` / ``-1``."""
    world, rank = _world(group)
    text_all, _ = all_gather_rows(text_local.contiguous(), group)
    dev = text_all.device
    nt, nv_local = text_all.shape[0], video_local.shape[0]
    if world > 1:
        nv = torch.tensor([nv_local], device=dev, dtype=torch.int64)
        sizes = [torch.zeros_like(nv) for _ in range(world)]
        dist.all_gather(sizes, nv, group=group)
        col_offset = int(sum(int(s.item()) for s in sizes[:rank]))
    else:
        col_offset = 0
    values = torch.full((nt, k), float("-inf"), device=dev, dtype=torch.float32)
    indices = torch.full((nt, k), -1, device=dev, dtype=torch.int64)
    kl = min(k, nv_local)
    if kl > 0:
        video_local = video_local.contiguous()
        for lo in range(0, nt, row_chunk):
            hi = min(nt, lo + row_chunk)
            scores = similarity_factory(text_all[lo:hi].contiguous(), video_local, terms).scores()
            v, i = topk_fn(scores, kl)
            values[lo:hi, :kl] = v
            indices[lo:hi, :kl] = i.to(torch.int64) + col_offset
    if world == 1:
        return values, indices
    # merge: candidates ordered by rank = by ascending column offset, so "first among equals" is the lowest video index
    all_v = [torch.empty_like(values) for _ in range(world)]
    all_i = [torch.empty_like(indices) for _ in range(world)]
    dist.all_gather(all_v, values, group=group)
    dist.all_gather(all_i, indices, group=group)
    cand_v = torch.cat(all_v, dim=1).contiguous()
    cand_i = torch.cat(all_i, dim=1)
    best_v, pos = topk_fn(cand_v, k)
    return best_v, torch.gather(cand_i, 1, pos.to(torch.int64))


def metrics_from_ranks(ranks: torch.Tensor, num_candidates: int) -> Dict[str, torch.Tensor]:
    """r1/r5/r10 (fp32 fractions) and mr (int64, lower median + 1) -- the keys of ``text_video_retrieval.py:21``."""
    recall, median, _ = ops.metrics_from_ranks(ranks.contiguous(), num_candidates)
    return {"r1": recall[0], "r5": recall[1], "r10": recall[2], "mr": median}


class TextVideoRetrievalModule(nn.Module):
    """Lightning-free twin of ``TextVideoRetrievalLightningModule`` for ``command=evaluate`` / ``predict``."""

    def __init__(self, encoder: VideoTextEncoder, init_temperature: float = 0.05, min_temperature: float = 0.001,
                 fit_temperature: bool = True, compute_rank: bool = False, group=None, similarity_terms: int = 3) -> None:
        super().__init__()
        self.encoder = encoder
        # video_text_module.py:32-34
        self.logit_scale = nn.Parameter(torch.tensor([-math.log(init_temperature)]), requires_grad=fit_temperature)
        self.max_logit_scale = nn.Parameter(torch.tensor([-math.log(min_temperature)]), requires_grad=False)
        self.group = group
        self.similarity_terms = similarity_terms
        self.metrics: Dict[str, Rank] = {"r1": Recall(), "r5": Recall(top_k=5), "r10": Recall(top_k=10),
                                         "mr": MedianRank()}
        if compute_rank:
            self.metrics["rank"] = Rank()
        self.logged: Dict[str, Any] = {}
        self._loss_sum = 0.0
        self._loss_weight = 0

    def forward(self, batch: MutableMapping[str, Any], _batch_idx: int = 0) -> TYPE_OUTPUT:
        batch.pop("video_id", None)  # video_text_module.py:38-41
        return self.encoder(**batch)

    def _step(self, batch, batch_idx: int = 0) -> TYPE_OUTPUT:
        return self(batch, batch_idx)

    def validation_step(self, batch, batch_idx: int = 0, dataloader_idx: Optional[int] = None):
        return self._step(batch, batch_idx), dataloader_idx

    def validation_step_end(self, output) -> TYPE_OUTPUT:
        (encoded_video, encoded_text), _ = output
        return self._validation_dataset_step_end((encoded_video, encoded_text))

    def _validation_dataset_step_end(self, output: TYPE_OUTPUT) -> TYPE_OUTPUT:
        # text_video_retrieval.py:44-58: gather the batch across ranks, scaled B x B scores, NCE loss
        encoded_video, _ = all_gather_rows(output[0].contiguous(), self.group)
        encoded_text, _ = all_gather_rows(output[1].contiguous(), self.group)
        batch_size = len(encoded_video)
        scale = float(self.logit_scale.exp())
        # `logit_scale * V @ T.T` == (logit_scale * V) @ T.T: rows = videos here
        scores = ops.Similarity(encoded_video, encoded_text, self.similarity_terms).scores(alpha=scale)
        loss = ops.nce_loss(scores)
        self._loss_sum += float(loss) * batch_size  # PL's batch-size weighted mean of `loss/val`
        self._loss_weight += batch_size
        return encoded_video, encoded_text

    def _validate_dataset(self, outputs: Sequence[TYPE_OUTPUT]) -> Dict[str, torch.Tensor]:
        # text_video_retrieval.py:67-83, fused: ranks straight from the embeddings
        encoded_videos, encoded_texts = (torch.cat(x) for x in zip(*outputs))
        ranks = retrieval_ranks(encoded_texts, encoded_videos, group=None, terms=self.similarity_terms)
        n_videos = encoded_videos.shape[0]
        result = {}
        for name, metric in self.metrics.items():
            metric.reset()
            metric.update_from_ranks(ranks, n_videos)
            result[name] = metric.compute()
        return result

    def validation_epoch_end(self, outputs: Sequence[TYPE_OUTPUT]) -> Dict[str, Any]:
        result = self._validate_dataset(outputs)
        if self._loss_weight:
            result["loss/val"] = self._loss_sum / self._loss_weight
        self._loss_sum, self._loss_weight = 0.0, 0
        self.logged = result
        return result

    def predict_step(self, batch, batch_idx: int = 0) -> Mapping[str, Any]:
        video_ids = batch.get("video_id")  # video_text_module.py:85-91
        encoded_video, encoded_text = self._step(batch, batch_idx)
        return {"encoded_videos": encoded_video, "encoded_texts": encoded_text, "video_ids": video_ids}
