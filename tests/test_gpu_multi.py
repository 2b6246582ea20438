"""N > 1 on real GPUs (NCCL): the column-sharded fused similarity + rank merge must give every rank the same integer
ranks as the single-GPU run on the gathered embeddings -- with uneven contiguous shards, no padding and no duplicated
samples.  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped on a one-GPU box, where the same protocol
is covered by the gloo tests on CPU (tests/test_distributed_gloo.py) and the kernel by tests/test_gpu_rank.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import oracle
        from fitclip_b200 import metrics_from_ranks, ops, retrieval_ranks, retrieval_topk, shard_bounds
        g = torch.Generator().manual_seed(0)
        v = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1)
        t = torch.nn.functional.normalize(v + 1.2 * torch.randn(n, 512, generator=g), dim=-1)
        v[5] = v[4]  # an exact duplicate video: a score tie that straddles nothing but exercises the tie rule
        lo, hi = shard_bounds(n, world, rank)
        ranks = retrieval_ranks(t[lo:hi].to(dev), v[lo:hi].to(dev))
        assert ranks.shape == (n,)
        # single-GPU run of the same kernel on the full matrices, and the reference ranking of its own scores
        full = ops.Similarity(t.to(dev), v.to(dev), 3)
        target = torch.arange(n, dtype=torch.int32, device=dev)
        single = full.counts(target, full.target_scores(target)).long()
        assert torch.equal(ranks, single), (rank, (ranks != single).nonzero().flatten().tolist()[:10])
        expect = oracle.ref_stable_rank(full.scores().cpu(), torch.arange(n))
        assert torch.equal(ranks.cpu(), expect)
        # distributed top-k (local top-k per slab -> all-gather -> merge) == top-k of the full score matrix, incl. the tie
        values, indices = retrieval_topk(t[lo:hi].to(dev), v[lo:hi].to(dev), k=10, row_chunk=300)
        full_scores = full.scores()
        order = torch.argsort(full_scores, dim=1, descending=True, stable=True)[:, :10]
        assert torch.equal(indices, order) and torch.equal(values, full_scores.gather(1, order))
        m = metrics_from_ranks(ranks, n)
        assert int(m["mr"]) == int(expect.median()) + 1
        torch.save(ranks.cpu(), os.path.join(out_dir, f"ranks_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4096])
def test_sharded_ranks_equal_single_gpu(tmp_path, n):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    world = min(world, 4) if n > 2000 else min(world, 3)  # 1001 over 3 ranks: shards 334 / 334 / 333
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    ranks = [torch.load(os.path.join(str(tmp_path), f"ranks_{r}.pt")) for r in range(world)]
    for r in ranks[1:]:
        assert torch.equal(r, ranks[0])  # identical on every rank
