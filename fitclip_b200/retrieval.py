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


NO_GROUP = False  # `group=NO_GROUP`: treat the inputs as complete -- no collective even if a process group exists


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`: ceil(n / world) per rank, last shards may be short/empty."""
    per = -(-n // world)
    return min(n, rank * per), min(n, (rank + 1) * per)


def _world(group) -> Tuple[int, int]:
    if group is not NO_GROUP and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _gather_sizes(counts: Sequence[int], device: torch.device, group) -> List[List[int]]:
    """ONE collective + ONE host read for all the per-rank counts a call needs: -> [rank][i]."""
    world, _ = _world(group)
    mine = torch.tensor(list(counts), device=device, dtype=torch.int64)
    out = torch.empty(world * len(counts), device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.view(world, len(counts)).tolist()


def all_gather_rows(x: torch.Tensor, group=None, sizes: Optional[Sequence[int]] = None,
                    total: Optional[int] = None) -> Tuple[torch.Tensor, List[int]]:
    """Concatenate per-rank row blocks of different heights. Returns the gathered tensor and the per-rank heights.

    ``total``: the blocks are the :func:`shard_bounds` shards of ``total`` rows -- heights are known analytically, the
    padded gather buffer IS the concatenation (no size exchange, no host synchronisation, no strip copies: every shard
    but the last non-empty one is full).  ``sizes``: heights known to the caller.  Neither: one size all-gather + one
    host read."""
    world, rank = _world(group)
    if world == 1:
        return x, [x.shape[0]]
    x = x.contiguous()
    if total is not None:
        per = -(-total // world)
        sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
        if x.shape[0] != sizes[rank]:
            raise ValueError(f"rank {rank} holds {x.shape[0]} rows, shard_bounds({total}, {world}, {rank}) has {sizes[rank]}")
        if x.shape[0] != per:
            padded = x.new_zeros((per, *x.shape[1:]))
            padded[:x.shape[0]] = x
            x = padded
        out = x.new_empty((world * per, *x.shape[1:]))
        dist.all_gather_into_tensor(out, x, group=group)
        return out[:total], sizes
    if sizes is None:
        sizes = [s[0] for s in _gather_sizes([x.shape[0]], x.device, group)]
    sizes = [int(v) for v in sizes]
    cap = max(sizes)
    if x.shape[0] != cap:
        padded = x.new_zeros((cap, *x.shape[1:]))
        padded[:x.shape[0]] = x
        x = padded
    out = x.new_empty((world * cap, *x.shape[1:]))  # (world * cap, ...): the shape gloo's all-gather accepts too
    dist.all_gather_into_tensor(out, x, group=group)
    if all(v == cap for v in sizes):
        return out, sizes
    return torch.cat([out[r * cap:r * cap + v] for r, v in enumerate(sizes)]), sizes


def retrieval_ranks(text_local: torch.Tensor, video_local: torch.Tensor, group=None, terms: int = 3,
                    target_local: Optional[torch.Tensor] = None,
                    similarity_factory: Callable[..., Any] = ops.Similarity,
                    totals: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """0-based rank of every query's target video among ALL videos, int64 ``(Nt_total,)``, same on every rank.

    ``text_local (nt_r, E)`` / ``video_local (nv_r, E)``: this rank's fp32 embeddings, contiguous shards in rank
    order.  ``target_local``: global video index per local query (default: query i <-> video i).
    ``totals = (Nt_total, Nv_total)``: the shards are the :func:`shard_bounds` shards of those totals, so every size and
    offset is known analytically and the call issues no size exchange and no host synchronisation (the launch queue
    keeps running); without it the two local counts travel in ONE small all-gather with one host read.
    ``group=NO_GROUP``: inputs are complete, no communication.
    ``similarity_factory`` exists so the host-side sharding logic can be exercised without a GPU (tests only)."""
    world, rank = _world(group)
    if world > 1 and totals is not None:
        text_all, _ = all_gather_rows(text_local, group, total=totals[0])
        col_offset, hi = shard_bounds(totals[1], world, rank)
        if video_local.shape[0] != hi - col_offset:
            raise ValueError(f"rank {rank} holds {video_local.shape[0]} videos, its shard of {totals[1]} has {hi - col_offset}")
        text_sizes = None
    elif world > 1:
        counts = _gather_sizes([text_local.shape[0], video_local.shape[0]], text_local.device, group)
        text_sizes = [c[0] for c in counts]
        text_all, _ = all_gather_rows(text_local, group, sizes=text_sizes)
        col_offset = sum(c[1] for c in counts[:rank])
    else:
        text_all, col_offset, text_sizes = text_local.contiguous(), 0, None
    if target_local is None:
        target = torch.arange(text_all.shape[0], device=text_all.device, dtype=torch.int32)
    elif world > 1 and totals is not None:
        target, _ = all_gather_rows(target_local.to(torch.int32), group, total=totals[0])
    else:
        target, _ = all_gather_rows(target_local.to(torch.int32).contiguous(), group, sizes=text_sizes)
    nt = text_all.shape[0]
    if video_local.shape[0] > 0:
        sim = similarity_factory(text_all.contiguous(), video_local.contiguous(), terms)
        tscore = sim.target_scores(target.contiguous(), col_offset)
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
    if world > 1:
        counts = _gather_sizes([text_local.shape[0], video_local.shape[0]], text_local.device, group)
        text_all, _ = all_gather_rows(text_local, group, sizes=[c[0] for c in counts])
        col_offset = sum(c[1] for c in counts[:rank])
    else:
        text_all, col_offset = text_local.contiguous(), 0
    dev = text_all.device
    nt, nv_local = text_all.shape[0], video_local.shape[0]
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
    all_v = values.new_empty((world * nt, k))
    all_i = indices.new_empty((world * nt, k))
    dist.all_gather_into_tensor(all_v, values, group=group)
    dist.all_gather_into_tensor(all_i, indices, group=group)
    cand_v = all_v.view(world, nt, k).permute(1, 0, 2).reshape(nt, world * k).contiguous()
    cand_i = all_i.view(world, nt, k).permute(1, 0, 2).reshape(nt, world * k)
    best_v, pos = topk_fn(cand_v, k)
    return best_v, torch.gather(cand_i, 1, pos.to(torch.int64))


def metrics_from_ranks(ranks: torch.Tensor, num_candidates: int) -> Dict[str, torch.Tensor]:
    """r1/r5/r10 (fp32 fractions) and mr (int64, lower median + 1) -- the keys of ``text_video_retrieval.py:21``."""
    recall, median, _ = ops.metrics_from_ranks(ranks.contiguous(), num_candidates)
    return {"r1": recall[0], "r5": recall[1], "r10": recall[2], "mr": median}


class TextVideoRetrievalModule(nn.Module):
    """Lightning-free twin of ``TextVideoRetrievalLightningModule`` for ``command=evaluate`` / ``predict``.

    Multi-process protocol: step outputs stay LOCAL (each rank keeps the embeddings of its own samples); the per-batch
    ``loss/val`` is computed on the batch gathered across ranks like the reference (``text_video_retrieval.py:44-58``),
    and ``validation_epoch_end`` hands the local shards to :func:`retrieval_ranks` with ``self.group``, which shards the
    similarity by video columns -- every rank ends up with the same global ranks and derives the metrics from them
    locally.  (The reference gathers in ``validation_step_end`` and evaluates the full matrix on every rank.)

    ``dataset_names`` with more than one name selects the reference's multi-dataset mode
    (``text_video_retrieval.py:28-37,84-93``): ``validation_step`` receives a ``dataloader_idx``, every metric is cloned
    per dataset under ``f"{metric}_{dataset}"``, ``loss/val_{dataset}`` is logged per dataset and ``validation_epoch_end``
    takes one output list per dataset."""

    def __init__(self, encoder: VideoTextEncoder, init_temperature: float = 0.05, min_temperature: float = 0.001,
                 fit_temperature: bool = True, compute_rank: bool = False, group=None, similarity_terms: int = 3,
                 dataset_names: Optional[Sequence[str]] = None,
                 similarity_factory: Callable[..., Any] = ops.Similarity,
                 nce_loss_fn: Callable[[torch.Tensor], torch.Tensor] = ops.nce_loss) -> None:
        super().__init__()
        self.encoder = encoder
        # video_text_module.py:32-34
        self.logit_scale = nn.Parameter(torch.tensor([-math.log(init_temperature)]), requires_grad=fit_temperature)
        self.max_logit_scale = nn.Parameter(torch.tensor([-math.log(min_temperature)]), requires_grad=False)
        self.group = group
        self.similarity_terms = similarity_terms
        self._similarity_factory = similarity_factory  # injectable so the host logic runs without a GPU (tests only)
        self._nce_loss = nce_loss_fn
        metrics: Dict[str, Rank] = {"r1": Recall(), "r5": Recall(top_k=5), "r10": Recall(top_k=10),
                                    "mr": MedianRank()}
        if compute_rank:
            metrics["rank"] = Rank()
        self.dataset_names = list(dataset_names) if dataset_names else None
        self.multiple_datasets = self.dataset_names is not None and len(self.dataset_names) > 1
        if self.multiple_datasets:
            assert all("_" not in name for name in self.dataset_names), \
                "Underscores in dataset names are problematic because of how we get their corresponding metrics."
            self.metrics: Dict[str, Rank] = {f"{name}_{dataset_name}": metric.clone()
                                             for dataset_name in self.dataset_names for name, metric in metrics.items()}
        else:
            self.metrics = metrics
        self.logged: Dict[str, Any] = {}
        self._loss: Dict[str, List[float]] = {}  # key -> [weighted sum, weight]

    def forward(self, batch: MutableMapping[str, Any], _batch_idx: int = 0) -> TYPE_OUTPUT:
        batch.pop("video_id", None)  # video_text_module.py:38-41
        return self.encoder(**batch)

    def _step(self, batch, batch_idx: int = 0) -> TYPE_OUTPUT:
        return self(batch, batch_idx)

    def validation_step(self, batch, batch_idx: int = 0, dataloader_idx: Optional[int] = None):
        return self._step(batch, batch_idx), dataloader_idx

    def validation_step_end(self, output) -> TYPE_OUTPUT:
        step_output, dataloader_idx = output
        assert self.multiple_datasets == (dataloader_idx is not None)  # text_video_retrieval.py:62-63
        dataset_name = self.dataset_names[dataloader_idx] if self.multiple_datasets else None
        return self._validation_dataset_step_end(step_output, dataset_name=dataset_name)

    def _validation_dataset_step_end(self, output: TYPE_OUTPUT, dataset_name: Optional[str] = None) -> TYPE_OUTPUT:
        # text_video_retrieval.py:44-58: gather the batch across ranks, scaled B x B scores, NCE loss
        assert len(output[0]) == len(output[1]), "retrieval batches pair every video with one caption"
        encoded_video, sizes = all_gather_rows(output[0].contiguous(), self.group)
        encoded_text, _ = all_gather_rows(output[1].contiguous(), self.group, sizes=sizes)  # same sizes: one exchange
        batch_size = len(encoded_video)
        scale = float(self.logit_scale.detach().exp())
        # `logit_scale * V @ T.T` == (logit_scale * V) @ T.T: rows = videos here
        scores = self._similarity_factory(encoded_video, encoded_text, self.similarity_terms).scores(alpha=scale)
        loss = self._nce_loss(scores)
        key = "loss/val" + ("" if dataset_name is None else f"_{dataset_name}")
        acc = self._loss.setdefault(key, [0.0, 0])  # PL's batch-size weighted mean of the logged value
        acc[0] += float(loss) * batch_size
        acc[1] += batch_size
        return output[0], output[1]  # local shards: retrieval_ranks does the sharded evaluation at epoch end

    def _validate_dataset(self, outputs: Sequence[TYPE_OUTPUT], dataset_name: Optional[str] = None
                          ) -> Dict[str, torch.Tensor]:
        # text_video_retrieval.py:67-83, fused: ranks straight from the embeddings
        assert self.multiple_datasets == (dataset_name is not None)
        encoded_videos, encoded_texts = (torch.cat(x) for x in zip(*outputs))
        ranks = retrieval_ranks(encoded_texts, encoded_videos, group=self.group, terms=self.similarity_terms,
                                similarity_factory=self._similarity_factory)
        n_videos = int(ranks.numel())  # query i <-> video i: the global gallery is as large as the global query set
        result = {}
        for name, metric in self.metrics.items():
            if not dataset_name or name.endswith(f"_{dataset_name}"):
                metric.reset()
                metric.update_from_ranks(ranks, n_videos)
                result[name] = metric.compute_local()  # the ranks are already global: no second reduction
        return result

    def validation_epoch_end(self, outputs) -> Dict[str, Any]:
        result: Dict[str, Any] = {}
        if self.multiple_datasets:
            for name, dataset_output in zip(self.dataset_names, outputs):  # text_video_retrieval.py:86-91
                result.update(self._validate_dataset(dataset_output, dataset_name=name))
        else:
            result.update(self._validate_dataset(outputs))
        for key, (total, weight) in self._loss.items():
            if weight:
                result[key] = total / weight
        self._loss = {}
        self.logged = result
        return result

    def predict_step(self, batch, batch_idx: int = 0) -> Mapping[str, Any]:
        video_ids = batch.get("video_id")  # video_text_module.py:85-91
        encoded_video, encoded_text = self._step(batch, batch_idx)
        return {"encoded_videos": encoded_video, "encoded_texts": encoded_text, "video_ids": video_ids}
