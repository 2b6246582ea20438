"""world_size-2 (and 3) gloo runs of the N>1 host logic on CPU: contiguous uneven shards, all-gather of text
embeddings, all-reduce of target scores and rank counts.  The similarity kernel itself needs a GPU, so the sharding
code is driven with a CPU stand-in that implements the same (target_scores, counts) contract with the oracle's tie
rule -- this file tests the protocol, tests/test_gpu_rank.py tests the kernel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fitclip_b200 import shard_bounds
from fitclip_b200.retrieval import all_gather_rows, retrieval_ranks, retrieval_topk


class CpuSimilarity:
    """Test double for fitclip_b200.ops.Similarity (same contract, plain torch fp32)."""

    def __init__(self, text_emb, video_emb, terms=3):
        self.s = text_emb @ video_emb.T
        self.nv = video_emb.shape[0]

    def target_scores(self, target, col_offset=0):
        local = target.long() - col_offset
        ok = (local >= 0) & (local < self.nv)
        out = torch.zeros(self.s.shape[0])
        out[ok] = self.s[ok, local[ok]]
        return out

    def scores(self, alpha=1.0):
        return alpha * self.s

    def counts(self, target, tscore, col_offset=0):
        gcol = torch.arange(self.nv).unsqueeze(0) + col_offset
        hit = (self.s > tscore.unsqueeze(1)) | ((self.s == tscore.unsqueeze(1)) & (gcol < target.long().unsqueeze(1)))
        return hit.sum(dim=1).to(torch.int32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ties, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        t = torch.nn.functional.normalize(torch.randn(n, 32, generator=g), dim=-1)
        v = torch.nn.functional.normalize(torch.randn(n, 32, generator=g), dim=-1)
        if ties:
            v[1::3] = v[0::3][: len(v[1::3])]  # duplicated videos -> exact score ties across shards
        lo, hi = shard_bounds(n, world, rank)
        ranks = retrieval_ranks(t[lo:hi], v[lo:hi], similarity_factory=CpuSimilarity)
        gathered, sizes = all_gather_rows(t[lo:hi])
        assert torch.equal(gathered, t) and sum(sizes) == n
        expect = oracle.ref_stable_rank(t @ v.T, torch.arange(n))
        assert torch.equal(ranks, expect), (rank, ranks.tolist(), expect.tolist())
        # analytic shard sizes (totals=): same result with no size exchange / host read
        ranks2 = retrieval_ranks(t[lo:hi], v[lo:hi], similarity_factory=CpuSimilarity, totals=(n, n))
        assert torch.equal(ranks2, expect), (rank, ranks2.tolist())
        gathered2, sizes2 = all_gather_rows(t[lo:hi], total=n)
        assert torch.equal(gathered2, t) and sizes2 == sizes
        # distributed top-k: local top-k per column slab, all-gather of the candidates, merge.  (Not on the tied case:
        # the CPU stand-in's per-slab matmuls differ from the full matmul in the last bit, which reorders exact ties; the
        # GPU kernel computes a score identically wherever its tile lies -- tests/test_gpu_multi.py covers ties.)
        if not ties:
            def cpu_topk(scores, k):
                order = torch.argsort(scores, dim=1, descending=True, stable=True)[:, :k]
                return scores.gather(1, order), order.to(torch.int32)
            k = min(5, n)
            values, indices = retrieval_topk(t[lo:hi], v[lo:hi], k=k, row_chunk=7, similarity_factory=CpuSimilarity,
                                             topk_fn=cpu_topk)
            full = t @ v.T
            order = torch.argsort(full, dim=1, descending=True, stable=True)[:, :k]
            assert torch.equal(indices, order), (rank, indices[:3].tolist(), order[:3].tolist())
            assert torch.allclose(values, full.gather(1, order), atol=1e-6)
        torch.save(ranks, os.path.join(out_dir, f"ranks_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,ties", [(2, 101, False), (2, 64, True), (3, 7, False), (2, 1, False)])
def test_sharded_ranks_equal_single_process(tmp_path, world, n, ties):
    mp.spawn(_worker, args=(world, _free_port(), n, ties, str(tmp_path)), nprocs=world, join=True)
    ranks = [torch.load(tmp_path / f"ranks_{r}.pt") for r in range(world)]
    assert all(torch.equal(ranks[0], r) for r in ranks)  # identical on every rank


def test_single_process_path_needs_no_process_group():
    g = torch.Generator().manual_seed(1)
    t, v = torch.randn(20, 16, generator=g), torch.randn(20, 16, generator=g)
    ranks = retrieval_ranks(t, v, similarity_factory=CpuSimilarity)
    assert torch.equal(ranks, oracle.ref_stable_rank(t @ v.T, torch.arange(20)))
    target = torch.randint(0, 20, (20,))
    ranks = retrieval_ranks(t, v, target_local=target, similarity_factory=CpuSimilarity)
    assert torch.equal(ranks, oracle.ref_stable_rank(t @ v.T, target))


def _metric_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fitclip_b200.metrics import Rank
        # "cat" state reduction (aligner/metrics.py:13) with uneven per-rank contributions, fed in two updates
        mine = torch.arange(10 * rank, 10 * rank + 3 + 2 * rank, dtype=torch.int64)
        m = Rank()
        m.update_from_ranks(mine[:2], 100)
        m.update_from_ranks(mine[2:], 100)
        allr = m.compute()  # Rank.compute is the plain concatenation: no kernel involved
        expect = torch.cat([torch.arange(10 * r, 10 * r + 3 + 2 * r, dtype=torch.int64) for r in range(world)])
        assert torch.equal(allr, expect), (rank, allr.tolist())
        # classification-style sharding: every rank holds ALL class columns and a slice of the query rows with explicit
        # targets; retrieval_ranks with target_local gathers queries and targets, columns are sharded
        g = torch.Generator().manual_seed(3)
        q = torch.randn(23, 16, generator=g)
        c = torch.randn(11, 16, generator=g)
        tgt = torch.randint(0, 11, (23,), generator=g)
        from fitclip_b200 import shard_bounds
        qlo, qhi = shard_bounds(23, world, rank)
        clo, chi = shard_bounds(11, world, rank)
        ranks = retrieval_ranks(q[qlo:qhi], c[clo:chi], target_local=tgt[qlo:qhi], similarity_factory=CpuSimilarity)
        assert torch.equal(ranks, oracle.ref_stable_rank(q @ c.T, tgt)), rank
        torch.save(allr, os.path.join(out_dir, f"m_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_metric_state_cat_and_explicit_targets(tmp_path):
    mp.spawn(_metric_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    outs = [torch.load(tmp_path / f"m_{r}.pt") for r in range(3)]
    assert all(torch.equal(o, outs[0]) for o in outs)


class _StubEncoder(torch.nn.Module):
    """Deterministic CPU stand-in for a VideoTextEncoder: 'video' / 'text' are already embeddings."""

    def forward(self, video, text):
        return video, text["input_ids"]


def _cpu_metrics_from_ranks(ranks, num_candidates):
    recall = torch.stack([(ranks < k).float().mean() for k in (1, 5, 10)])
    return recall, ranks.median() + 1, ranks.float().mean() + 1


def _module_worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fitclip_b200 import TextVideoRetrievalModule, ops
        ops.metrics_from_ranks = _cpu_metrics_from_ranks  # the metric reduction kernel needs a GPU; same contract
        g = torch.Generator().manual_seed(5)
        t = torch.nn.functional.normalize(torch.randn(n, 16, generator=g), dim=-1)
        v = torch.nn.functional.normalize(t + 0.7 * torch.randn(n, 16, generator=g), dim=-1)
        module = TextVideoRetrievalModule(_StubEncoder(), compute_rank=True, similarity_factory=CpuSimilarity,
                                          nce_loss_fn=oracle.ref_nce_loss)
        lo, hi = shard_bounds(n, world, rank)
        outputs = []
        for b in range(lo, hi, 5):  # batches of 5 of this rank's shard; every rank runs the same number of steps
            e = min(hi, b + 5)
            batch = {"video": v[b:e], "text": {"input_ids": t[b:e]}, "video_id": list(range(b, e))}
            outputs.append(module.validation_step_end(module.validation_step(batch, len(outputs))))
        result = module.validation_epoch_end(outputs)
        torch.save({k: torch.as_tensor(x) for k, x in result.items()}, os.path.join(out_dir, f"mod_{world}_{rank}.pt"))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_retrieval_module_two_ranks_equals_single_process(tmp_path):
    """with a process group, the module must not gather twice -- world=2 metrics == single-process
    metrics (r1/r5/r10/mr and the full rank list), identical on both ranks."""
    n = 40  # two shards of 20 = four batches of 5 per rank
    mp.spawn(_module_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    _module_worker(0, 1, 0, n, str(tmp_path))
    single = torch.load(tmp_path / "mod_1_0.pt")
    for r in range(2):
        got = torch.load(tmp_path / f"mod_2_{r}.pt")
        for k in ("r1", "r5", "r10", "mr", "rank"):
            assert torch.equal(got[k], single[k]), (r, k, got[k], single[k])
    assert 0.0 < float(single["r1"]) < 1.0  # a non-degenerate case


def test_retrieval_module_multiple_datasets():
    """``dataset_names`` with two names (text_video_retrieval.py:28-37, 60-65, 84-93): ``validation_step`` carries a
    ``dataloader_idx``, every metric is cloned per dataset under ``{metric}_{dataset}``, ``loss/val_{dataset}`` is logged
    per dataset, ``validation_epoch_end`` takes one output list per dataset -- each dataset's numbers equal a
    single-dataset module fed that dataset alone."""
    from fitclip_b200 import TextVideoRetrievalModule, ops
    ops_backup = ops.metrics_from_ranks
    ops.metrics_from_ranks = _cpu_metrics_from_ranks
    try:
        g = torch.Generator().manual_seed(11)
        data = {}
        for name, n in (("msrvtt", 25), ("didemo", 15)):
            t = torch.nn.functional.normalize(torch.randn(n, 16, generator=g), dim=-1)
            v = torch.nn.functional.normalize(t + 0.8 * torch.randn(n, 16, generator=g), dim=-1)
            data[name] = (v, t)
        kw = dict(similarity_factory=CpuSimilarity, nce_loss_fn=oracle.ref_nce_loss)
        multi = TextVideoRetrievalModule(_StubEncoder(), dataset_names=list(data), **kw)
        assert multi.multiple_datasets and set(multi.metrics) == {f"{m}_{d}" for d in data for m in ("r1", "r5", "r10", "mr")}
        outputs = []
        for idx, (name, (v, t)) in enumerate(data.items()):
            outs = []
            for b in range(0, len(v), 5):
                batch = {"video": v[b:b + 5], "text": {"input_ids": t[b:b + 5]}}
                outs.append(multi.validation_step_end(multi.validation_step(batch, len(outs), dataloader_idx=idx)))
            outputs.append(outs)
        result = multi.validation_epoch_end(outputs)
        for name, (v, t) in data.items():
            single = TextVideoRetrievalModule(_StubEncoder(), **kw)
            outs = [single.validation_step_end(single.validation_step(
                {"video": v[b:b + 5], "text": {"input_ids": t[b:b + 5]}}, 0)) for b in range(0, len(v), 5)]
            expect = single.validation_epoch_end(outs)
            for m in ("r1", "r5", "r10", "mr"):
                assert torch.equal(torch.as_tensor(result[f"{m}_{name}"]), torch.as_tensor(expect[m])), (name, m)
            assert abs(result[f"loss/val_{name}"] - expect["loss/val"]) < 1e-6
        with pytest.raises(AssertionError):  # a dataloader_idx is mandatory in multi-dataset mode (:62-63)
            multi.validation_step_end(multi.validation_step({"video": v[:2], "text": {"input_ids": t[:2]}}, 0))
    finally:
        ops.metrics_from_ranks = ops_backup
