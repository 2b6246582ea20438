"""bench.py's own inputs (CPU): the CUDA arm's weights / captions come from the package's initialiser and from bench.py, the CPU
oracle of `cpu_baseline` / `--impl reference` loads the same state dict; the single-GPU record of the 100k-video leg carries a
fingerprint of the kernel sources."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402


def test_bench_weights_load_into_the_oracle_and_are_trained_like():
    sd = bench.synthetic_weights(0)
    model = oracle.CLIP(**oracle.clip_ref.VIT_B_16)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    g = sd["visual.transformer.resblocks.3.ln_1.weight"]
    assert 0.2 <= float(g.min()) and float(g.max()) <= 3.0 and float(g.std()) > 0.5   # no identity LayerNorm anywhere
    assert float(sd["transformer.resblocks.0.attn.in_proj_bias"].abs().max()) > 0
    assert float(sd["visual.ln_post.bias"].abs().max()) > 0
    again = bench.synthetic_weights(0)
    assert all(torch.equal(sd[k], again[k]) for k in sd)                               # every rank builds the same model
    assert not torch.equal(sd["text_projection"], bench.synthetic_weights(1)["text_projection"])


def test_bench_captions_are_dense_and_framed():
    ids = bench.synthetic_tokens(16, 4321)
    assert ids.shape == (16, bench.CTX) and ids.dtype == torch.int32
    assert (ids[:, 0] == 49406).all() and (ids[:, -1] == 49407).all()
    assert int(ids[:, 1:-1].min()) >= 1 and int(ids[:, 1:-1].max()) < 49406            # EOT is the row maximum, at the end
    assert (ids.argmax(dim=1) == bench.CTX - 1).all()


def test_single_gpu_record_names_the_kernel_sources_it_was_made_with():
    h = bench.kernel_sources_sha256()
    assert len(h) == 64 and h == bench.kernel_sources_sha256()
    with open(os.path.join(ROOT, "profiles", "r2_webvid_1gpu.json")) as f:
        rec = json.load(f)
    assert rec["videos"] == 100_000 and rec["n_gpus"] == 1 and len(rec["kernel_sources_sha256"]) == 64
    assert set(rec["metrics"]) == {"r1", "r5", "r10", "mr"}
