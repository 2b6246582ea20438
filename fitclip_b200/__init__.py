"""fitclip_b200 -- B200-native (sm_100a) implementation of FitCLIP's evaluation hot path behind the reference's
``VideoTextEncoder`` plugin interface.  See DESIGN.md / INTEGRATION.md."""
from .api import VideoEncoder, VideoTextEncoder  # noqa: F401
from .classification import VideoTextClassificationModule  # noqa: F401
from .encoder import B200Clip, B200ClipVideoTextEncoder, load_clip_model  # noqa: F401
from .metrics import Accuracy, MeanRank, MedianRank, Rank, Recall  # noqa: F401
from .retrieval import (TextVideoRetrievalModule, metrics_from_ranks, retrieval_ranks, retrieval_topk,  # noqa: F401
                        shard_bounds)
from .slip_encoder import B200SlipClip, B200SlipVideoTextEncoder, load_slip_model  # noqa: F401
from .teacher_student import TeacherStudentScoringModule  # noqa: F401
from .training import ClipTrainer, TeacherStudentTrainingModule, VideoTextTrainingModule  # noqa: F401
from .wise import wise, wise_state_dict  # noqa: F401

__version__ = "0.1.0"
