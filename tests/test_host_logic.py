"""Host-side mirror of the reference plugin interface: everything that must work without a GPU."""
import copy
import pickle

import pytest
import torch

import oracle
from fitclip_b200 import B200Clip, B200ClipVideoTextEncoder, VideoTextEncoder, _lib, load_clip_model, shard_bounds
from fitclip_b200 import tokenizer
from fitclip_b200.frame_sampler import RandomFromUniformIntervalsFrameSampler, UniformFrameSampler

TINY = dict(embed_dim=64, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=16,
            context_length=77, vocab_size=64, transformer_width=64, transformer_heads=1, transformer_layers=2)


@pytest.fixture(scope="module")
def ref_model():
    return oracle.clip_vit_b_16(seed=0, **TINY)


def test_parameter_names_match_the_openai_layout(ref_model):
    enc = B200ClipVideoTextEncoder(ref_model.state_dict())
    ref = oracle.RefClipVideoTextEncoder(copy.deepcopy(ref_model))
    assert isinstance(enc, VideoTextEncoder) and isinstance(enc, torch.nn.Module)
    assert [k for k, _ in enc.named_parameters()] == [k for k, _ in ref.named_parameters()]
    assert all(k.startswith("model.") for k in enc.state_dict())
    assert "model.logit_scale" not in enc.state_dict()  # clip_video_text_encoder.py:75-77
    assert enc.model.visual.input_resolution == 32 and enc.model.dtype == torch.float32
    # names pinned by config/trainer/callbacks/clip_freeze_text.yaml:25-29
    for prefix in ("model.token_embedding.", "model.positional_embedding", "model.transformer.", "model.ln_final.",
                   "model.text_projection"):
        assert any(k.startswith(prefix) for k in enc.state_dict()), prefix


def test_accepts_module_state_dict_or_b200clip(ref_model):
    a = B200ClipVideoTextEncoder(ref_model)
    b = B200ClipVideoTextEncoder(ref_model.state_dict())
    c = B200ClipVideoTextEncoder(B200Clip(ref_model.state_dict()), num_frames=8)
    for k in a.state_dict():
        assert torch.equal(a.state_dict()[k], b.state_dict()[k]) and torch.equal(a.state_dict()[k], c.state_dict()[k])
    assert c.num_frames == 8 and a.num_frames == 4
    assert a.model.config == {**TINY}


def test_deepcopy_and_strict_load_state_dict_round_trip(ref_model):
    enc = B200ClipVideoTextEncoder(ref_model.state_dict())
    clone = copy.deepcopy(enc)  # aligner/wise.py:21
    other = oracle.RefClipVideoTextEncoder(oracle.clip_vit_b_16(seed=1, **TINY))
    clone.load_state_dict(other.state_dict(), strict=True)  # same keys as the reference wrapper
    assert torch.equal(clone.state_dict()["model.text_projection"], other.state_dict()["model.text_projection"])
    assert not torch.equal(enc.state_dict()["model.text_projection"], clone.state_dict()["model.text_projection"])
    with pytest.raises(RuntimeError):
        clone.load_state_dict({"model.text_projection": torch.zeros(64, 64)}, strict=True)
    pickle.loads(pickle.dumps(enc.model._engine))  # the native handle never travels


def test_load_clip_model_from_file_without_logit_scale(tmp_path, ref_model):
    sd = {k: v for k, v in ref_model.state_dict().items() if k != "logit_scale"}  # student checkpoints lack it
    path = tmp_path / "student.pt"
    torch.save(sd, path)
    model = load_clip_model(str(path))
    assert isinstance(model, B200Clip) and next(model.parameters()).device.type == "cpu"
    assert not hasattr(model, "logit_scale")
    with pytest.raises(FileNotFoundError):
        load_clip_model("ViT-B/16")  # names/URLs need the network


def test_no_cpu_fallback(ref_model):
    enc = B200ClipVideoTextEncoder(ref_model.state_dict())
    with pytest.raises(_lib.FitclipError):
        enc.encode_video(torch.zeros(1, 2, 3, 32, 32))
    with pytest.raises(_lib.FitclipError):
        enc.encode_text({"input_ids": torch.zeros(1, 77, dtype=torch.int32)})
    from fitclip_b200 import ops
    with pytest.raises(_lib.FitclipError):
        ops.rank_from_scores(torch.zeros(2, 2), torch.zeros(2, dtype=torch.long))


def test_hooks(ref_model):
    enc = B200ClipVideoTextEncoder(ref_model.state_dict(), num_frames=4)
    assert enc.should_pad_batch is True
    t = torch.zeros(2, 3)
    assert enc.to_bchw(t) is t
    assert isinstance(enc.get_eval_frame_sampler(), UniformFrameSampler)
    assert isinstance(enc.get_train_frame_sampler(), RandomFromUniformIntervalsFrameSampler)
    frames = torch.randint(0, 255, (3, 48, 64, 3), dtype=torch.uint8)  # (T, H, W, C) as decoded
    out = enc.get_eval_transform(torch.float32)(frames)
    assert out.shape == (3, 3, 32, 32) and out.dtype == torch.float32
    out = enc.get_train_transform(torch.float32)(frames)
    assert out.shape == (3, 3, 32, 32)
    pickle.dumps(enc.get_tokenizer())  # must be picklable for DataLoader workers
    pickle.dumps(enc.get_eval_frame_sampler())
    video = torch.zeros(1, 3, 4, 4)
    restored = enc.denormalize_video_tensor(video.clone())
    assert restored.dtype == torch.uint8 and restored[0, 0, 0, 0] == int(0.48145466 * 255)


def test_uniform_frame_sampler_midpoints():
    s = UniformFrameSampler(4)
    assert [int(i) for i in s(0, 99, 30.0)] == [12, 36, 62, 86]  # midpoints of linspace(0, 99, 5) as ints
    assert [int(i) for i in s(0, 1, 30.0)] == [0, 0]  # 2 frames only: int ticks [0, 0, 1], round-half-even midpoints
    r = RandomFromUniformIntervalsFrameSampler(4)
    torch.manual_seed(0)
    idx = [int(i) for i in r(0, 99, 30.0)]
    ticks = [0, 24, 49, 74, 99]
    assert all(a <= i <= b for i, a, b in zip(idx, ticks[:-1], ticks[1:]))


def test_token_padding_and_truncation_rule():
    ids = tokenizer.pad_tokens([[5, 6, 7], list(range(100, 300)), []])
    assert ids.dtype == torch.int32 and ids.shape == (3, 77)
    assert ids[0, :6].tolist() == [49406, 5, 6, 7, 49407, 0]
    assert ids[1, 0] == 49406 and ids[1, 76] == 49407 and ids[1, 75] == 174  # cut to 77, EOT forced last
    assert ids[2, :3].tolist() == [49406, 49407, 0]
    assert ids.argmax(dim=-1).tolist() == [4, 76, 1]
    with pytest.raises(RuntimeError):
        tokenizer.pad_tokens([list(range(100, 300))], truncate=False)


def test_shard_bounds_are_contiguous_and_cover_everything():
    for n, w in ((1000, 8), (1001, 8), (5, 8), (100000, 3), (0, 2)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert all(hi - lo <= -(-n // w) for lo, hi in spans)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfitclip_b200.so")
    with pytest.raises(_lib.FitclipError, match="no fallback"):
        _lib.load()


def test_unsupported_layouts_are_named_in_the_error():
    import oracle
    from fitclip_b200 import B200Clip, _lib
    sd = oracle.clip_vit_b_16(seed=0, vision_layers=1, transformer_layers=1, vision_width=64, transformer_width=64,
                              transformer_heads=1, embed_dim=64, image_resolution=32, context_length=8,
                              vocab_size=64).state_dict()
    resnet = dict(sd)
    resnet["visual.layer1.0.conv1.weight"] = torch.zeros(64, 64, 1, 1)  # clip_rn50.yaml and friends
    with pytest.raises(_lib.FitclipError, match="ModifiedResNet"):
        B200Clip(resnet)
    slip = oracle.slip_clip_vit_b_16(seed=0, img_size=32, vision_width=64, vision_layers=1, vision_heads=1, embed_dim=64,
                                     context_length=8, vocab_size=64, transformer_width=64, transformer_heads=1,
                                     transformer_layers=1).state_dict()
    with pytest.raises(_lib.FitclipError, match="SLIP-layout"):
        B200Clip(slip)
