"""CPU oracle for the FitCLIP evaluation hot path.

TEST INFRASTRUCTURE ONLY. Nothing under ``fitclip_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

Parity status: **partly pinned against the reference's own code, partly unpinned**.  The reference repository holds no
golden vectors, known-answer tests or fixtures for this path, and the arithmetic of its vision
tower and of ``Recall`` / ``Accuracy`` lives in un-vendored third-party packages (openai/CLIP@b46f5ac, torchmetrics 0.9)
that are not importable here.  What IS in the reference tree has been executed in the build container (third-party
imports stubbed, no reference file copied: ``tests/golden/make_reference_golden.py``) and its outputs are frozen in
``tests/golden/reference_outputs.pt``: the ``ClipVideoTextEncoder`` wrapper (``encode_video`` / ``encode_text``), the
in-tree twin of the CLIP text tower (``aligner/encoder/slip.py:350-480``), ``wise`` / ``wise_state_dict``, ``nce_loss`` /
``teacher_student_nce_loss``, ``Rank`` / ``MedianRank``, the eval frame sampler, and the evaluation flows of
``TextVideoRetrievalLightningModule``, ``VideoTextClassificationLightningModule`` and ``TeacherStudentLightningModule``
(on a stub ``pl.LightningModule``) -- ``tests/test_reference_golden.py`` checks this oracle (CPU) and the CUDA path (GPU)
against them.  The training step (``train_ref.py``) is pinned the same way: ``tests/golden/make_reference_training_golden.py``
runs the reference's ``training_step`` / ``training_step_end`` + autograd + AdamW and ``tests/test_reference_training_golden.py``
checks the oracle and the CUDA path against the frozen loss / gradients / update.  The rest is pinned against (a) independent
implementations in the image -- ``transformers.CLIPModel`` for both towers (tests/test_oracle_clip.py), torchvision for
the eval transform (tests/test_oracle_preprocess.py) -- and (b) seeded golden vectors frozen under ``tests/golden/`` by
``tests/golden/make_golden.py``; **the vision tower and the torchmetrics definitions stay "parity unpinned"** in the
sense of the task statement.  The SLIP-layout models (row f4; ``slip_ref.py``): timm's ``VisionTransformer`` is likewise
third-party and absent -- restated, pinned against ``transformers.ViTModel`` (tests/test_oracle_slip.py); the reference's own
``slip.CLIP`` + ``SlipVideoTextEncoder`` around it were run in the build container
(``tests/golden/make_reference_slip_golden.py`` -> ``tests/golden/reference_slip.pt``).
"""
from .clip_ref import CLIP, build_model, clip_vit_b_16, perturb_trained_like, tokenize_synthetic  # noqa: F401
from .encoder_ref import RefClipVideoTextEncoder  # noqa: F401
from .metrics_ref import (ref_accuracy_at_k, ref_median_rank, ref_rank, ref_recall_at_k,  # noqa: F401
                          ref_retrieval_metrics, ref_stable_rank)
from .wise_ref import ref_wise, ref_wise_state_dict  # noqa: F401
from .preprocess_ref import ref_eval_transform, ref_resized_size  # noqa: F401
from .loss_ref import ref_nce_loss, ref_teacher_student_nce_loss  # noqa: F401
from .bf16_emulation import bf16_stream_model  # noqa: F401
from .train_ref import ref_training_loss, ref_training_step  # noqa: F401
from .slip_ref import (RefSlipVideoTextEncoder, SlipClip, TimmVisionTransformer, slip_clip_vit_b_16)  # noqa: F401
