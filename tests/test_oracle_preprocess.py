"""The pre-processing oracle against the torchvision pipeline the reference builds
(aligner/encoder/clip_video_text_encoder.py:124-133), and the CPU hook the encoder returns."""
import pytest
import torch

import oracle
from fitclip_b200.transforms import eval_transform

MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


@pytest.mark.parametrize("h,w", [(240, 320), (360, 202), (224, 224), (256, 256), (113, 400), (480, 270)])
def test_oracle_matches_torchvision_pipeline(h, w):
    g = torch.Generator().manual_seed(h * 1000 + w)
    video = torch.randint(0, 256, (3, h, w, 3), dtype=torch.uint8, generator=g)
    ours = oracle.ref_eval_transform(video, 224, MEAN, STD)
    hook = eval_transform(224, torch.float32, MEAN, STD)(video)
    assert ours.shape == hook.shape == (3, 3, 224, 224)
    assert torch.allclose(ours, hook, atol=1e-6, rtol=0), (ours - hook).abs().max().item()


def test_resized_size_rule():
    # new_short, new_long = size, int(size * long / short): truncation, not rounding
    assert oracle.ref_resized_size(240, 320, 224) == (224, 298)
    assert oracle.ref_resized_size(360, 202, 224) == (399, 224)
    assert oracle.ref_resized_size(224, 224, 224) == (224, 224)


def test_eval_hook_does_not_antialias():
    """torchvision >= 0.17 antialiases by default; the reference (0.12) does not -- the hook must spell it out."""
    g = torch.Generator().manual_seed(3)
    video = torch.randint(0, 256, (1, 448, 448, 3), dtype=torch.uint8, generator=g)
    hook = eval_transform(224, torch.float32, MEAN, STD)(video)
    x = video.permute(0, 3, 1, 2).float() / 255
    plain = torch.nn.functional.interpolate(x, size=(224, 224), mode="bicubic", align_corners=False)
    m = torch.tensor(MEAN).view(1, 3, 1, 1)
    s = torch.tensor(STD).view(1, 3, 1, 1)
    assert torch.allclose(hook, (plain - m) / s, atol=1e-6)


@pytest.mark.parametrize("h,w", [(240, 320), (360, 202), (113, 400)])
def test_bilinear_oracle_matches_the_slip_wrapper_transform(h, w):
    """``SlipVideoTextEncoder.get_eval_transform`` (slip_video_text_encoder.py:78-87): ``T.Resize(size)`` with its default
    interpolation (bilinear, no antialias for tensors in the pinned torchvision), ImageNet statistics."""
    from torchvision.transforms import InterpolationMode
    from fitclip_b200.slip_encoder import IMAGENET_MEAN, IMAGENET_STD
    g = torch.Generator().manual_seed(h + w)
    video = torch.randint(0, 256, (2, h, w, 3), dtype=torch.uint8, generator=g)
    ours = oracle.ref_eval_transform(video, 224, IMAGENET_MEAN, IMAGENET_STD, interpolation="bilinear")
    hook = eval_transform(224, torch.float32, IMAGENET_MEAN, IMAGENET_STD, interpolation=InterpolationMode.BILINEAR)(video)
    assert torch.allclose(ours, hook, atol=1e-6, rtol=0), (ours - hook).abs().max().item()
    assert not torch.allclose(ours, oracle.ref_eval_transform(video, 224, IMAGENET_MEAN, IMAGENET_STD), atol=1e-3)
