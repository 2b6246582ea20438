"""BASELINE.json configs[3] on one GPU, in two measured parts (the full 100k x 8-frame encode is ~29 PFLOP, half a minute
of pure compute, so the encode rate is measured on a slice and the 100k x 100k similarity + rank in full):

  1. encode rate at 8 frames per video: `videos` videos x 8 frames (frames generated on the device chunk by chunk --
     100k x 8 fp32 frames would be 482 GB) + as many 77-token captions;
  2. similarity + ranks + R@k / MdR over a 100 000 x 100 000 gallery (never materialised), plus the top-10 lists.

Prints one JSON line (copied to profiles/)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from fitclip_b200 import B200ClipVideoTextEncoder, metrics_from_ranks, retrieval_ranks, retrieval_topk  # noqa: E402

dev = torch.device("cuda:0")
videos = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
T, chunk = 8, 250
enc = B200ClipVideoTextEncoder(oracle.clip_vit_b_16(seed=0).state_dict(), num_frames=T).to(dev)
ids = oracle.tokenize_synthetic(videos, 77, seed=4321).to(dev)
g = torch.Generator(device=dev).manual_seed(1234)
buf = torch.randn(chunk, T, 3, 224, 224, device=dev, generator=g)


def encode_all():
    out = []
    for lo in range(0, videos, chunk):
        buf.normal_(generator=g)  # fresh frames every chunk: nothing is cached
        out.append(enc.encode_video(buf[:min(chunk, videos - lo)]))
    return torch.cat(out), enc.encode_text({"input_ids": ids})


with torch.inference_mode():
    encode_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v, t = encode_all()
    e1.record()
    torch.cuda.synchronize()
    enc_ms = e0.elapsed_time(e1)

    n = 100_000
    gv = torch.nn.functional.normalize(torch.randn(n, 512, device=dev, generator=g), dim=-1)
    gt = torch.nn.functional.normalize(gv + 1.5 * torch.randn(n, 512, device=dev, generator=g), dim=-1)
    retrieval_ranks(gt[:1024].contiguous(), gv[:1024].contiguous())
    torch.cuda.synchronize()
    e0.record()
    ranks = retrieval_ranks(gt, gv)
    m = metrics_from_ranks(ranks, n)
    e1.record()
    torch.cuda.synchronize()
    rank_ms = e0.elapsed_time(e1)
    e0.record()
    topv, topi = retrieval_topk(gt, gv, k=10)
    e1.record()
    torch.cuda.synchronize()
    topk_ms = e0.elapsed_time(e1)

flop_per_video = 8 * 35_126_906_880 + 5_959_540_736
rate = videos / (enc_ms * 1e-3)
full_s = n / rate + rank_ms * 1e-3
print(json.dumps({
    "workload": "webvid_scale: 8 frames/video, ViT-B/16, 1 GPU", "encode_videos": videos, "encode_ms": round(enc_ms, 1),
    "encode_videos_per_s": round(rate, 1), "encode_tflops": round(rate * flop_per_video / 1e12, 1),
    "gallery": n, "sim_rank_metrics_ms": round(rank_ms, 1), "sim_rank_tflops": round(2 * 2.0 * n * n * 1536 / rank_ms / 1e9, 1),
    "top10_ms": round(topk_ms, 1), "metrics": {k: float(x) for k, x in m.items()},
    "projected_100k_eval_s": round(full_s, 1), "projected_videos_per_s": round(n / full_s, 1),
    "projected_roofline_frac_of_1649.9": round(n * flop_per_video / full_s / 1649.9e12, 3)}))
