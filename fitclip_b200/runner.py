"""Lightning/Hydra-free runner for the reference's command line (``aligner/__main__.py:27-93``, ``aligner/cli.py:81-150``):

    python -m aligner command=evaluate encoder=clip_vit_b_16 data=synthetic_msrvtt
    python -m aligner command=evaluate encoder=wise +encoder@encoder.model1=clip_vit_b_16 \
        +encoder@encoder.model2=clip_vit_b_16 encoder.model2.model.seed=1 data=synthetic_ucf101
    python -m aligner command=predict encoder=clip_vit_b_16 data=synthetic_msrvtt output_path=predictions.pt
    python -m aligner --config-name teacher_student_trainer command=train +encoder@encoder.student=clip_vit_b_16 \
        +encoder@encoder.teacher=clip_vit_b_16 encoder.teacher.model.seed=1 data=synthetic_teacher_student

It implements the subset of Hydra the reference relies on for evaluation: a root config with a
``defaults`` list, ``group=name`` selection with per-file ``defaults`` chains, ``+group@package=name`` placement,
dotted ``key=value`` / ``+key=value`` overrides, ``${key}`` interpolation of top-level scalars and recursive
``_target_`` instantiation.  The loop reproduces the hook order of PL 1.6's evaluation loop: ``on_validation_start``,
``validation_step`` -> ``validation_step_end`` per batch, ``validation_epoch_end`` once, and prints the logged keys.
Real datasets / video decoding are out of scope (no network, no decord): ``data=synthetic_*`` feeds tensors of the
shapes and dtypes the reference's data modules emit."""
from __future__ import annotations

import copy
import importlib
import math
import json
import os
import re
import sys
from typing import Any, Dict, Iterator, List, Mapping, Optional, Sequence

import torch
import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config")


# ------------------------------------------------------------------------------------------------- config resolution
def _load_yaml(path: str) -> Dict[str, Any]:
    with open(path) as f:
        return yaml.safe_load(f) or {}


def _merge(dst: Dict[str, Any], src: Mapping[str, Any]) -> Dict[str, Any]:
    for k, v in src.items():
        if isinstance(v, Mapping) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def load_group(group: str, name: str, config_dir: str = CONFIG_DIR) -> Dict[str, Any]:
    """``config/<group>/<name>.yaml`` with its own ``defaults`` chain resolved (entries of the same group first,
    ``_self_`` last unless placed explicitly; ``- key: null`` placeholders are skipped)."""
    cfg = _load_yaml(os.path.join(config_dir, group, f"{name}.yaml"))
    defaults = cfg.pop("defaults", [])
    merged: Dict[str, Any] = {}
    self_done = False
    for entry in defaults:
        if entry == "_self_":
            _merge(merged, cfg)
            self_done = True
        elif isinstance(entry, str):
            _merge(merged, load_group(group, entry, config_dir))
        # mappings such as {model1: null} are optional slots filled from the command line
    if not self_done:
        _merge(merged, cfg)
    return merged


def _set_path(cfg: Dict[str, Any], dotted: str, value: Any) -> None:
    *parents, leaf = dotted.split(".")
    node = cfg
    for p in parents:
        if not isinstance(node.get(p), dict):
            node[p] = {}
        node = node[p]
    if isinstance(value, Mapping) and isinstance(node.get(leaf), dict):
        _merge(node[leaf], value)
    else:
        node[leaf] = value


def _interpolate(node: Any, root: Mapping[str, Any]) -> Any:
    if isinstance(node, dict):
        return {k: _interpolate(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_interpolate(v, root) for v in node]
    if isinstance(node, str):
        m = re.fullmatch(r"\$\{(\w+)\}", node)
        if m:
            return root[m.group(1)]
        return re.sub(r"\$\{(\w+)\}", lambda mm: str(root[mm.group(1)]), node)
    return node


def compose(overrides: Sequence[str], config_dir: str = CONFIG_DIR, config_name: str = "trainer") -> Dict[str, Any]:
    cfg = _load_yaml(os.path.join(config_dir, f"{config_name}.yaml"))
    defaults = cfg.pop("defaults", [])
    groups = [next(iter(d)) for d in defaults if isinstance(d, Mapping)]
    # a root config may build on another root config (`defaults: [trainer, _self_]`, config/teacher_student_trainer.yaml)
    for entry in defaults:
        if isinstance(entry, str) and entry != "_self_":
            base = _load_yaml(os.path.join(config_dir, f"{entry}.yaml"))
            groups += [next(iter(d)) for d in base.pop("defaults", []) if isinstance(d, Mapping)]
            cfg = _merge(base, cfg)
    values: List[tuple] = []
    for ov in overrides:
        key, _, raw = ov.partition("=")
        key = key.lstrip("+")
        if "@" in key:  # +group@package.path=name
            group, package = key.split("@", 1)
            _set_path(cfg, package, load_group(group, raw, config_dir))
        elif key in groups and os.path.exists(os.path.join(config_dir, key, f"{raw}.yaml")):
            cfg[key] = load_group(key, raw, config_dir)
        else:
            values.append((key, yaml.safe_load(raw)))
    for key, value in values:  # plain value overrides win over group contents
        _set_path(cfg, key, value)
    cfg = _interpolate(cfg, cfg)
    def _unset(node: Any, path: str) -> List[str]:
        if node == "???":
            return [path]
        if isinstance(node, Mapping) and path == "encoder":  # encoder maps: encoder.student / encoder.teacher
            return [p for k, v in node.items() if v == "???" for p in [f"{path}.{k}"]]
        return []

    missing = [p for k, v in cfg.items() for p in _unset(v, k)]
    if missing:
        raise ValueError(f"Missing mandatory value(s): {', '.join(missing)} (e.g. command=evaluate encoder=clip_vit_b_16 "
                         f"data=synthetic_msrvtt)")
    return cfg


def instantiate(node: Any, **extra: Any) -> Any:
    """``hydra.utils.instantiate``: children first, then import ``_target_`` and call it with the remaining keys."""
    if isinstance(node, Mapping):
        if "_target_" in node:
            kwargs = {k: instantiate(v) for k, v in node.items() if not k.startswith("_")}
            kwargs.update(extra)
            unresolved = [k for k, v in kwargs.items() if v == "???"]
            if unresolved:
                raise ValueError(f"{node['_target_']}: missing mandatory value(s) {unresolved}")
            module, _, attr = node["_target_"].rpartition(".")
            return getattr(importlib.import_module(module), attr)(**kwargs)
        return {k: instantiate(v) for k, v in node.items()}
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    return node


# ------------------------------------------------------------------------------------------------- synthetic plug-ins
def random_init_clip(seed: int = 0, **kwargs: int):
    """``clip.model.CLIP(**kwargs)`` with its published initialisation (config/encoder/clip_from_scratch_vit_b_16.yaml);
    returns a :class:`fitclip_b200.B200Clip` holding the parameters."""
    from . import B200Clip
    from ._init import init_clip_state_dict
    return B200Clip(init_clip_state_dict(seed=seed, **kwargs))


def random_init_slip(seed: int = 0, vision_heads: Optional[int] = None, **kwargs: int):
    """``slip.CLIP_VITB16()``-style random init (config/encoder/slip_from_scratch_vit_b_16.yaml): a
    :class:`fitclip_b200.B200SlipClip` with timm-ViT image tower parameters under the SLIP checkpoint names.
    ``vision_heads``: 12 for the ViT-S/16 variant (heads of 32, config/encoder/slip_vit_s_16.yaml)."""
    from . import B200SlipClip
    from ._init import init_slip_state_dict
    return B200SlipClip(init_slip_state_dict(seed=seed, **kwargs), vision_heads=vision_heads)


class SyntheticRetrievalData:
    """Batches shaped like ``VideoTextDataModule`` output: ``{"video": (B,T,3,R,R) fp32, "text": {"input_ids": (B,77)},
    "video_id": [...]}`` -- N(0,1) frames generated on the device, ``[SOT] + random ids + [EOT]`` captions."""

    def __init__(self, encoder, num_videos: int = 1000, batch_size: int = 32, caption_length: int = 77,
                 seed: int = 1234) -> None:
        self.encoder, self.num_videos, self.batch_size = encoder, num_videos, batch_size
        self.caption_length, self.seed = caption_length, seed

    def _geometry(self):
        enc = self.encoder
        model = enc.model
        return enc.num_frames, model.visual.input_resolution, model.context_length, model.vocab_size

    def val_batches(self, device: torch.device) -> Iterator[Dict[str, Any]]:
        frames, res, ctx, vocab = self._geometry()
        for i, lo in enumerate(range(0, self.num_videos, self.batch_size)):
            b = min(self.batch_size, self.num_videos - lo)
            g = torch.Generator(device=device).manual_seed(self.seed + i)
            video = torch.randn(b, frames, 3, res, res, device=device, generator=g)
            gc = torch.Generator().manual_seed(self.seed + 7919 * (i + 1))
            ids = torch.zeros(b, ctx, dtype=torch.int32)
            n = min(self.caption_length, ctx)
            ids[:, :n] = torch.randint(1, vocab - 2, (b, n), generator=gc, dtype=torch.int32)
            ids[:, 0] = vocab - 2
            ids[:, n - 1] = vocab - 1
            yield {"video": video, "text": {"input_ids": ids.to(device)},
                   "video_id": [f"video{lo + j}" for j in range(b)]}

    def train_batches(self, device: torch.device, steps: int) -> Iterator[Dict[str, Any]]:
        """``steps`` training batches of ``batch_size`` (video, caption) pairs (a fresh seed per step)."""
        saved = self.num_videos, self.seed
        try:
            for i in range(steps):
                self.num_videos, self.seed = self.batch_size, saved[1] + 104_729 * (i + 1)
                yield next(iter(self.val_batches(device)))
        finally:
            self.num_videos, self.seed = saved


class SyntheticClassificationData(SyntheticRetrievalData):
    """Batches shaped like ``VideoClassificationDataModule`` output (``target = (name, id)``) plus the label prompts."""

    def __init__(self, encoder, num_videos: int = 3783, num_frames: int = 8, num_labels: int = 101,
                 num_templates: int = 48, batch_size: int = 32, seed: int = 1234) -> None:
        super().__init__(encoder, num_videos, batch_size, 16, seed)
        self.num_frames, self.num_labels, self.num_templates = num_frames, num_labels, num_templates
        self.categories = [f"class{i}" for i in range(num_labels)]
        self.templates = [f"template{t} of {{}}" for t in range(num_templates)]

    def tokenized_prompts(self) -> Dict[str, torch.Tensor]:
        _, _, ctx, vocab = self._geometry()
        g = torch.Generator().manual_seed(self.seed + 1)
        n = self.num_labels * self.num_templates
        lengths = torch.randint(6, 20, (n,), generator=g)
        ids = torch.zeros(n, ctx, dtype=torch.int32)
        body = torch.randint(1, vocab - 2, (n, ctx), generator=g, dtype=torch.int32)
        for i in range(n):
            k = int(lengths[i])
            ids[i, :k] = body[i, :k]
            ids[i, 0], ids[i, k - 1] = vocab - 2, vocab - 1
        return {"input_ids": ids}

    def val_batches(self, device: torch.device) -> Iterator[Dict[str, Any]]:
        _, res, _, _ = self._geometry()
        gl = torch.Generator().manual_seed(7)
        labels = torch.randint(0, self.num_labels, (self.num_videos,), generator=gl)
        for i, lo in enumerate(range(0, self.num_videos, self.batch_size)):
            b = min(self.batch_size, self.num_videos - lo)
            g = torch.Generator(device=device).manual_seed(self.seed + i)
            video = torch.randn(b, self.num_frames, 3, res, res, device=device, generator=g)
            y = labels[lo:lo + b].to(device)
            yield {"video": video, "target": ([self.categories[int(t)] for t in y], y),
                   "video_id": [f"video{lo + j}" for j in range(b)]}


class SyntheticTeacherStudentData:
    """Training batches with the keys the reference's data modules emit for an encoder map
    (``aligner/data/video_dataset.py:40-56``, ``tokenizer_collate.py:84-87``): ``video_student`` / ``video_teacher``,
    ``text_student`` / ``text_teacher`` and ``dataset`` (one name per sample, grouped: labelled first)."""

    def __init__(self, encoder: Mapping[str, Any], batch_size: int = 512, labeled_fraction: float = 0.5,
                 caption_length: int = 77, seed: int = 1234) -> None:
        self.encoder, self.batch_size, self.labeled_fraction = encoder, batch_size, labeled_fraction
        self.caption_length, self.seed = caption_length, seed

    def train_batches(self, device: torch.device, steps: int) -> Iterator[Dict[str, Any]]:
        enc = self.encoder["student"]
        frames, res = enc.num_frames, enc.model.visual.input_resolution
        ctx, vocab = enc.model.context_length, enc.model.vocab_size
        n_lab = int(round(self.batch_size * self.labeled_fraction))
        names = ["labeled"] * n_lab + ["unlabeled"] * (self.batch_size - n_lab)
        for i in range(steps):
            g = torch.Generator(device=device).manual_seed(self.seed + i)
            video = torch.randn(self.batch_size, frames, 3, res, res, device=device, generator=g)
            n = min(self.caption_length, ctx)
            ids = torch.zeros(self.batch_size, ctx, dtype=torch.int32, device=device)
            ids[:, :n] = torch.randint(1, vocab - 2, (self.batch_size, n), device=device, generator=g, dtype=torch.int32)
            ids[:, 0] = vocab - 2
            ids[:, n - 1] = vocab - 1
            yield {"video_student": video, "video_teacher": video, "text_student": {"input_ids": ids},
                   "text_teacher": {"input_ids": ids}, "dataset": names}


# ------------------------------------------------------------------------------------------------- commands
def _to_float(v: Any) -> Any:
    if isinstance(v, torch.Tensor):
        return v.tolist() if v.numel() > 1 else v.item()
    return v


def evaluate(cfg: Mapping[str, Any], device: Optional[torch.device] = None) -> Dict[str, Any]:
    from . import VideoTextClassificationModule
    device = device or torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(cfg.get("seed", 42))  # init_cli: seed_everything (aligner/cli.py:43-50)
    encoder = instantiate(cfg["encoder"]).to(device)
    data = instantiate(cfg["data"], encoder=encoder)
    if isinstance(data, SyntheticClassificationData):  # aligner/cli.py:110-115 switches the model class the same way
        model = VideoTextClassificationModule(encoder, data.categories, data.templates,
                                              tokenized_labels=data.tokenized_prompts())
    else:
        model = instantiate(cfg["model"], encoder=encoder).to(device)
    command = cfg["command"]
    with torch.inference_mode():
        if hasattr(model, "on_validation_start"):
            model.on_validation_start()
        if command == "predict":
            outs = [model.predict_step(batch, i) for i, batch in enumerate(data.val_batches(device))]
            result = {k: (torch.cat([o[k] for o in outs]).cpu() if isinstance(outs[0][k], torch.Tensor)
                          else [x for o in outs for x in o[k]]) for k in outs[0]}
            torch.save(result, cfg.get("output_path", "predictions.pt"))  # aligner/__main__.py:70-91
            return {"saved": cfg.get("output_path", "predictions.pt"), "count": len(result["video_ids"])}
        outputs = []
        for i, batch in enumerate(data.val_batches(device)):
            out = model.validation_step(batch, i)
            if hasattr(model, "validation_step_end"):
                out = model.validation_step_end(out)
            outputs.append(out)
        result = model.validation_epoch_end(outputs)
    if hasattr(encoder, "model") and hasattr(encoder.model, "check_inputs"):
        encoder.model.check_inputs()
    return {k: _to_float(v) for k, v in result.items()}


def train(cfg: Mapping[str, Any], device: Optional[torch.device] = None) -> Dict[str, Any]:
    """``command=train`` with an encoder map (``aligner/__main__.py:49-62``: ``trainer.fit``): the teacher-student
    training loop, ``trainer.max_steps`` steps, logging ``loss/train`` per step like ``training_step_end``
    (``aligner/teacher_student.py:176-183``)."""
    if "_target_" in cfg["encoder"]:  # a single encoder: VideoTextLightningModule's NCE training (video_text_module.py:25-97)
        return _train_single(cfg, device)
    if not isinstance(cfg["encoder"], Mapping) or set(cfg["encoder"]) != {"student", "teacher"}:
        raise ValueError("command=train needs one encoder (NCE fine-tuning) or an encoder map: --config-name "
                         "teacher_student_trainer with +encoder@encoder.student=... +encoder@encoder.teacher=...")
    device = device or torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(cfg.get("seed", 42))
    encoders = {k: v.to(device) for k, v in instantiate(cfg["encoder"]).items()}
    data = instantiate(cfg["data"], encoder=encoders)
    opt = cfg.get("optimizer", {})
    extra = {}
    if cfg.get("prompts"):  # aligner/cli.py:117-121: a text file, one prompt per non-empty line
        with open(cfg["prompts"]) as file:
            extra["prompts"] = [line.strip() for line in file if line.strip()]
    if cfg.get("trainer", {}).get("gradient_clip_val") is not None:  # config/trainer.yaml:39
        extra["gradient_clip_val"] = float(cfg["trainer"]["gradient_clip_val"])
    model = instantiate(cfg["model"], encoder=encoders["student"], teacher=encoders["teacher"],
                        lr=float(opt.get("lr", 3e-6)), weight_decay=float(opt.get("weight_decay", 1e-2)), **extra)
    steps = int(cfg.get("trainer", {}).get("max_steps", 10))
    losses = []
    for i, batch in enumerate(data.train_batches(device, steps)):
        losses.append(model.training_step(batch, i))
    losses = [float(x) for x in losses]  # one device read-back at the end
    model.trainer.check_inputs()  # the student's training forward (its own err_flag, not the eval engine's)
    if hasattr(encoders["teacher"].model, "check_inputs"):
        encoders["teacher"].model.check_inputs()
    return {"loss/train": losses[-1], "step": len(losses), "losses": losses}


def _train_single(cfg: Mapping[str, Any], device: Optional[torch.device] = None) -> Dict[str, Any]:
    """``command=train`` with one encoder: the reference fits its ``TextVideoRetrievalLightningModule`` -- i.e.
    ``VideoTextLightningModule.training_step`` / ``training_step_end`` (NCE on the gathered batch, logit scale trained when
    ``model.fit_temperature``) -- for ``trainer.max_steps`` steps."""
    from .training import VideoTextTrainingModule
    device = device or torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(cfg.get("seed", 42))
    encoder = instantiate(cfg["encoder"]).to(device)
    data = instantiate(cfg["data"], encoder=encoder)
    opt, mcfg = cfg.get("optimizer", {}), cfg.get("model", {})
    model = VideoTextTrainingModule(encoder, init_temperature=float(mcfg.get("init_temperature", 0.05)),
                                    min_temperature=float(mcfg.get("min_temperature", 0.001)),
                                    fit_temperature=bool(mcfg.get("fit_temperature", True)),
                                    lr=float(opt.get("lr", 3e-6)), weight_decay=float(opt.get("weight_decay", 1e-2)),
                                    gradient_clip_val=cfg.get("trainer", {}).get("gradient_clip_val"))
    steps = int(cfg.get("trainer", {}).get("max_steps", 10))
    losses = [model.training_step(batch, i) for i, batch in enumerate(data.train_batches(device, steps))]
    losses = [float(x) for x in losses]  # one device read-back at the end
    model.trainer.check_inputs()
    return {"loss/train": losses[-1], "step": len(losses), "losses": losses, "temperature": math.exp(-model.logit_scale)}


def main(argv: Sequence[str]) -> int:
    argv = list(argv)
    config_name = "trainer"
    for i, a in enumerate(argv):  # hydra's --config-name / -cn
        if a in ("--config-name", "-cn"):
            config_name = argv[i + 1]
            del argv[i:i + 2]
            break
        if a.startswith("--config-name="):
            config_name = a.split("=", 1)[1]
            del argv[i]
            break
    cfg = compose(argv, config_name=config_name)
    command = cfg["command"]
    if command not in ("evaluate", "validate", "test", "predict", "train"):
        raise ValueError(f"command={command} is not implemented (supported: evaluate | validate | test | predict | train)")
    result = train(cfg) if command == "train" else evaluate(cfg)
    if not cfg.get("silent"):
        width = max(len(k) for k in result)
        for k, v in result.items():
            if not (isinstance(v, list) and len(v) > 20):
                print(f"{k:<{width}}  {v}")
    print(json.dumps({k: v for k, v in result.items() if not isinstance(v, list)}))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
