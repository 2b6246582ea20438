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


def _train_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import oracle
        from fitclip_b200 import B200ClipVideoTextEncoder
        from fitclip_b200.training import TeacherStudentTrainingModule
        geom = dict(vision_layers=2, transformer_layers=2, image_resolution=64, context_length=24, vocab_size=512)
        enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **geom).state_dict(), num_frames=2).to(dev)
        teach = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1, **geom).state_dict(), num_frames=2).to(dev)
        module = TeacherStudentTrainingModule(enc, teach, lr=1e-4)
        n = 16
        g = torch.Generator().manual_seed(0)
        video = torch.randn(n, 2, 3, 64, 64, generator=g)
        ids = oracle.tokenize_synthetic(n, (5, 24), seed=1, context_length=24, vocab_size=512)
        per = n // world
        sl = slice(rank * per, (rank + 1) * per)
        batch = {"video_student": video[sl].to(dev), "video_teacher": video[sl].to(dev),
                 "text_student": {"input_ids": ids[sl].to(dev)}, "text_teacher": {"input_ids": ids[sl].to(dev)}}
        loss = module.training_step(batch, 0, optimize=True)  # gathers embeddings, all-reduces the flat gradient
        torch.save({"loss": loss.cpu(), "grad": module.trainer.grad.cpu(), "flat": module.trainer.flat.cpu()},
                   os.path.join(out_dir, f"train_{rank}.pt"))
        if rank == 0:  # the same step on the whole batch on one GPU
            enc1 = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0, **geom).state_dict(), num_frames=2).to(dev)
            single = TeacherStudentTrainingModule(enc1, teach, lr=1e-4, group=False)
            full = {"video_student": video.to(dev), "video_teacher": video.to(dev),
                    "text_student": {"input_ids": ids.to(dev)}, "text_teacher": {"input_ids": ids.to(dev)}}
            loss1 = single.training_step(full, 0, optimize=False)
            torch.save({"loss": loss1.cpu(), "grad": single.trainer.grad.cpu()}, os.path.join(out_dir, "train_single.pt"))
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_equals_single_gpu(tmp_path):
    """2 ranks x 8 samples (embeddings all-gathered over NCCL, one all-reduce of the flat gradient buffer) == 16 samples
    on one GPU: same loss, same gradient (bf16 activations, atomics: cosine >= 0.999), replicas bit-identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    mp.spawn(_train_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(os.path.join(str(tmp_path), f"train_{r}.pt")) for r in range(2))
    single = torch.load(os.path.join(str(tmp_path), "train_single.pt"))
    assert torch.equal(a["grad"], b["grad"]) and torch.equal(a["flat"], b["flat"])
    assert abs(float(a["loss"]) - float(single["loss"])) <= 1e-3 * abs(float(single["loss"]))
    cos = torch.nn.functional.cosine_similarity(a["grad"], single["grad"], dim=0)
    assert float(cos) >= 0.999, float(cos)
