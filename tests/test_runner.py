"""`python -m aligner ...` shim: config composition (CPU) and an end-to-end evaluate / predict run (GPU)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from fitclip_b200 import runner
from fitclip_b200._init import init_clip_state_dict, init_slip_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["encoder.model.vision_layers=1", "encoder.model.transformer_layers=1"]


def test_compose_encoder_chain_and_overrides():
    cfg = runner.compose(["command=evaluate", "encoder=clip_vit_b_16", "data=synthetic_msrvtt", "data.num_videos=64",
                          "+encoder.num_frames=2"])
    assert cfg["encoder"]["_target_"] == "fitclip_b200.B200ClipVideoTextEncoder"  # from clip.yaml through the chain
    assert cfg["encoder"]["model"]["vision_width"] == 768 and cfg["encoder"]["model"]["transformer_heads"] == 8
    assert cfg["encoder"]["num_frames"] == 2 and cfg["data"]["num_videos"] == 64
    assert cfg["model"]["init_temperature"] == 0.015 and cfg["model"]["fit_temperature"] is False


def test_compose_wise_with_package_placement():
    cfg = runner.compose(["command=evaluate", "encoder=wise", "+encoder@encoder.model1=clip_vit_b_16",
                          "+encoder@encoder.model2=clip_vit_b_16", "encoder.model2.model.seed=1",
                          "encoder.weight_for_2=0.5", "data=synthetic_msrvtt"])
    enc = cfg["encoder"]
    assert enc["_target_"] == "fitclip_b200.wise.wise" and enc["weight_for_2"] == 0.5
    assert enc["model1"]["model"]["seed"] == 0 and enc["model2"]["model"]["seed"] == 1


def test_missing_mandatory_values_are_reported():
    with pytest.raises(ValueError, match="Missing mandatory"):
        runner.compose(["command=evaluate"])
    with pytest.raises(ValueError, match="encoder map"):  # neither one encoder nor a student / teacher map
        runner.train({"command": "train", "encoder": {"a": {"_target_": "x"}}, "data": {}})


def test_instantiate_is_recursive_and_rejects_unfilled_slots():
    obj = runner.instantiate({"_target_": "collections.OrderedDict", "a": {"_target_": "datetime.timedelta", "days": 7}})
    assert obj["a"].days == 7
    with pytest.raises(ValueError):
        runner.instantiate({"_target_": "builtins.dict", "model1": "???"})


def test_random_init_statistics_match_the_published_init():
    sd = init_clip_state_dict(seed=0, vision_layers=1, transformer_layers=2)
    assert len(sd) == 1 + 5 + 12 * 2 + 8 + 12 * 1 and sd["visual.positional_embedding"].shape == (197, 768)
    w = sd["transformer.resblocks.0.attn.in_proj_weight"]
    assert abs(w.std().item() - 512 ** -0.5) < 2e-3
    assert abs(sd["transformer.resblocks.1.mlp.c_proj.weight"].std().item() - (512 ** -0.5) * (4 ** -0.5)) < 1e-3
    assert abs(sd["token_embedding.weight"].std().item() - 0.02) < 1e-3
    v = sd["visual.transformer.resblocks.0.attn.in_proj_weight"]
    assert v.abs().max().item() <= (6.0 / (4 * 768)) ** 0.5 + 1e-6  # xavier-uniform bound
    assert torch.count_nonzero(sd["visual.transformer.resblocks.0.attn.in_proj_bias"]) == 0
    # deterministic per seed, different across seeds
    assert torch.equal(sd["text_projection"], init_clip_state_dict(seed=0, vision_layers=1, transformer_layers=2)["text_projection"])
    assert not torch.equal(sd["text_projection"], init_clip_state_dict(seed=1, vision_layers=1, transformer_layers=2)["text_projection"])


def test_compose_slip_encoder_chain_and_slip_init():
    cfg = runner.compose(["command=evaluate", "encoder=slip_from_scratch_vit_b_16", "data=synthetic_msrvtt"])
    assert cfg["encoder"]["_target_"] == "fitclip_b200.B200SlipVideoTextEncoder"  # from slip.yaml through the chain
    assert cfg["encoder"]["model"]["_target_"] == "fitclip_b200.runner.random_init_slip"
    with pytest.raises(ValueError, match="missing mandatory"):  # slip_from_pretrained needs a local checkpoint path
        runner.instantiate(runner.compose(["command=evaluate", "encoder=slip_from_pretrained",
                                           "data=synthetic_msrvtt"])["encoder"])
    # timm's init (slip.py:595-600 builds timm.create_model('vit_base_patch16_224')): truncated normal 0.02, zero biases
    sd = init_slip_state_dict(seed=0, vision_layers=1, transformer_layers=1)
    assert sd["visual.pos_embed"].shape == (1, 197, 768) and sd["visual.cls_token"].shape == (1, 1, 768)
    w = sd["visual.blocks.0.mlp.fc1.weight"]
    assert abs(w.std().item() - 0.02) < 1e-3 and w.abs().max().item() <= 2.0  # trunc_normal_'s cut is +-2 ABSOLUTE
    assert torch.count_nonzero(sd["visual.blocks.0.attn.qkv.bias"]) == 0
    assert abs(sd["image_projection"].std().item() - 768 ** -0.5) < 1e-3
    model = runner.random_init_slip(seed=0, vision_layers=1, transformer_layers=1)
    assert model.config["vision_tower"] == 1 and model.visual.input_resolution == 224


@pytest.mark.parametrize("name,expect", [
    ("clip_vit_b_32", dict(vision_patch_size=32, vision_width=768)),
    ("clip_vit_l_14", dict(vision_patch_size=14, vision_width=1024, vision_layers=24, embed_dim=768, transformer_width=768)),
    ("clip_vit_l_14_336px", dict(image_resolution=336, vision_patch_size=14, vision_width=1024)),
    ("slip_vit_b_16", dict(vision_width=768, vision_layers=12)),
    ("slip_vit_l_16", dict(vision_width=1024, vision_layers=24)),
    ("slip_vit_s_16", dict(vision_width=384, vision_heads=12))])
def test_named_geometry_configs_resolve(name, expect):
    """The reference's per-checkpoint encoder configs (config/encoder/<name>.yaml) resolve here to the same geometry,
    random-initialised offline."""
    cfg = runner.compose(["command=evaluate", f"encoder={name}", "data=synthetic_msrvtt"])["encoder"]
    assert cfg["_target_"].endswith("B200SlipVideoTextEncoder" if name.startswith("slip") else "B200ClipVideoTextEncoder")
    for k, v in expect.items():
        assert cfg["model"][k] == v, (name, k, cfg["model"].get(k))


def test_slip_vit_s_16_config_builds_padded_heads():
    cfg = runner.compose(["command=evaluate", "encoder=slip_vit_s_16", "data=synthetic_msrvtt",
                          "encoder.model.vision_layers=1", "encoder.model.transformer_layers=1"])
    enc = runner.instantiate(cfg["encoder"])
    assert enc.model.vision_heads == 12 and enc.model.config["vision_attn_width"] == 768


@pytest.mark.gpu
def test_evaluate_command_on_a_slip_layout_encoder():
    cfg = runner.compose(["command=evaluate", "encoder=slip_from_scratch_vit_b_16", "data=synthetic_msrvtt",
                          "data.num_videos=48", *TINY])
    result = runner.evaluate(cfg)
    assert set(result) == {"r1", "r5", "r10", "mr", "loss/val"} and 1 <= int(result["mr"]) <= 48


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["clip_vit_b_32", "slip_vit_s_16"])
def test_evaluate_command_on_named_geometries(name):
    cfg = runner.compose(["command=evaluate", f"encoder={name}", "data=synthetic_msrvtt", "data.num_videos=40", *TINY])
    result = runner.evaluate(cfg)
    assert set(result) == {"r1", "r5", "r10", "mr", "loss/val"} and 1 <= int(result["mr"]) <= 40


@pytest.mark.gpu
def test_evaluate_and_predict_commands(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-m", "aligner", "command=evaluate", "encoder=clip_vit_b_16",
                          "data=synthetic_msrvtt", "data.num_videos=80", *TINY], cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    result = json.loads(out.stdout.strip().splitlines()[-1])
    assert set(result) == {"r1", "r5", "r10", "mr", "loss/val"}
    assert 0.0 <= result["r10"] <= 1.0 and 1 <= result["mr"] <= 80
    pred = tmp_path / "predictions.pt"
    out = subprocess.run([sys.executable, "-m", "aligner", "command=predict", "encoder=clip_vit_b_16",
                          "data=synthetic_msrvtt", "data.num_videos=40", f"output_path={pred}", *TINY], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    saved = torch.load(pred)
    assert saved["encoded_videos"].shape == (40, 512) and saved["encoded_texts"].shape == (40, 512)
    assert len(saved["video_ids"]) == 40  # the reference's predictions.pt keys (aligner/video_text_module.py:85-91)


@pytest.mark.gpu
def test_evaluate_wise_classification():
    cfg = runner.compose(["command=evaluate", "encoder=wise", "+encoder@encoder.model1=clip_vit_b_16",
                          "+encoder@encoder.model2=clip_vit_b_16", "encoder.model2.model.seed=1",
                          "encoder.weight_for_2=0.5", "data=synthetic_ucf101", "data.num_videos=24", "data.num_frames=2",
                          "data.num_labels=7", "data.num_templates=3",
                          "encoder.model1.model.vision_layers=1", "encoder.model1.model.transformer_layers=1",
                          "encoder.model2.model.vision_layers=1", "encoder.model2.model.transformer_layers=1"])
    result = runner.evaluate(cfg)  # the models are built on the CPU like the reference's; the lerp itself runs on the GPU
    assert set(result) == {"a1", "a5", "mr"} and 1 <= int(result["mr"]) <= 7


def test_compose_teacher_student_root_config():
    cfg = runner.compose(["command=train", "+encoder@encoder.student=clip_vit_b_16", "+encoder@encoder.teacher=clip_vit_b_16",
                          "encoder.teacher.model.seed=1", "data=synthetic_teacher_student", "data.batch_size=8",
                          "trainer.max_steps=3"], config_name="teacher_student_trainer")
    assert cfg["model"]["_target_"].endswith("TeacherStudentTrainingModule") and cfg["trainer"]["max_steps"] == 3
    assert cfg["encoder"]["teacher"]["model"]["seed"] == 1 and cfg["encoder"]["student"]["model"]["seed"] == 0
    assert cfg["optimizer"]["lr"] == 3.0e-6 and cfg["seed"] == 42  # optimizer of config/trainer.yaml:22-24; base config merged
    with pytest.raises(ValueError, match="encoder.teacher"):
        runner.compose(["command=train", "+encoder@encoder.student=clip_vit_b_16", "data=synthetic_teacher_student"],
                       config_name="teacher_student_trainer")


@pytest.mark.gpu
def test_train_command():
    env = dict(os.environ, PYTHONPATH=ROOT)
    tiny = [f"encoder.{who}.model.{k}=1" for who in ("student", "teacher") for k in ("vision_layers", "transformer_layers")]
    out = subprocess.run([sys.executable, "-m", "aligner", "--config-name", "teacher_student_trainer", "command=train",
                          "+encoder@encoder.student=clip_vit_b_16", "+encoder@encoder.teacher=clip_vit_b_16",
                          "encoder.teacher.model.seed=1", "data=synthetic_teacher_student", "data.batch_size=16",
                          "trainer.max_steps=4", "optimizer.lr=1e-4", *tiny], cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    result = json.loads(out.stdout.strip().splitlines()[-1])
    assert result["step"] == 4 and result["loss/train"] == result["loss/train"]  # finite


@pytest.mark.gpu
def test_train_command_with_a_prompts_file(tmp_path):
    """``prompts=<file>`` (aligner/cli.py:117-121 -> teacher_student.py:47,104-120): the unlabelled captions are replaced by
    the file's lines, tokenised by the in-tree byte-pair tokenizer on the (synthetic, tests-only) merges file."""
    prompts = tmp_path / "prompts.txt"
    prompts.write_text("a video of a man playing guitar\n\n a woman is cooking pasta \npeople are dancing\n")
    env = dict(os.environ, PYTHONPATH=ROOT,
               FITCLIP_BPE_VOCAB=os.path.join(ROOT, "tests", "golden", "bpe_synthetic_vocab.txt.gz"))
    tiny = [f"encoder.{who}.model.{k}=1" for who in ("student", "teacher") for k in ("vision_layers", "transformer_layers")]
    out = subprocess.run([sys.executable, "-m", "aligner", "--config-name", "teacher_student_trainer", "command=train",
                          "+encoder@encoder.student=clip_vit_b_16", "+encoder@encoder.teacher=clip_vit_b_16",
                          "encoder.teacher.model.seed=1", "data=synthetic_teacher_student", "data.batch_size=16",
                          "trainer.max_steps=3", "optimizer.lr=1e-4", f"+prompts={prompts}", *tiny], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    result = json.loads(out.stdout.strip().splitlines()[-1])
    assert result["step"] == 3 and result["loss/train"] == result["loss/train"]


@pytest.mark.gpu
def test_train_command_with_a_single_encoder():
    """``command=train`` with one encoder (the reference fits its TextVideoRetrievalLightningModule: NCE on the batch,
    logit scale trained and clamped, video_text_module.py:25-97)."""
    cfg = runner.compose(["command=train", "encoder=clip_vit_b_16", "data=synthetic_msrvtt", "data.batch_size=16",
                          "+trainer.max_steps=4", "+optimizer.lr=1e-4", "model.fit_temperature=true", *TINY])
    result = runner.train(cfg)
    assert result["step"] == 4 and result["loss/train"] == result["loss/train"]
    assert all(x == x and x < 1e4 for x in result["losses"])  # finite on every step
    assert 0.001 <= result["temperature"] <= 0.05 + 1e-6
