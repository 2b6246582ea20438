"""CPU oracle for the FitCLIP evaluation hot path.

TEST INFRASTRUCTURE ONLY. Nothing under ``fitclip_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

Parity status: **parity unpinned by the reference** -- the reference repository holds no golden vectors,
known-answer tests or fixtures for this path and its arithmetic lives in
un-vendored third-party packages (openai/CLIP@b46f5ac, torchmetrics 0.9) that are not importable here.
The restatement is therefore pinned against (a) an independent implementation that *is* in the image,
``transformers.CLIPModel`` (tests/test_oracle_clip.py), (b) the in-tree structural twin
``aligner/encoder/slip.py:350-480`` that it follows line by line, and (c) seeded golden vectors frozen under
``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
from .clip_ref import CLIP, build_model, clip_vit_b_16, tokenize_synthetic  # noqa: F401
from .encoder_ref import RefClipVideoTextEncoder  # noqa: F401
from .metrics_ref import (ref_accuracy_at_k, ref_median_rank, ref_rank, ref_recall_at_k,  # noqa: F401
                          ref_retrieval_metrics, ref_stable_rank)
from .wise_ref import ref_wise, ref_wise_state_dict  # noqa: F401
from .preprocess_ref import ref_eval_transform, ref_resized_size  # noqa: F401
from .loss_ref import ref_nce_loss, ref_teacher_student_nce_loss  # noqa: F401
