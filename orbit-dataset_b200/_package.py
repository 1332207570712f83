"""orbit_b200: B200-native (sm_100a) implementation of the ORBIT few-shot recogniser's episodic hot path."""
from .lib import OrbitError, load as load_library  # noqa: F401
from .few_shot_recognisers import (FewShotRecogniser, MultiStepFewShotRecogniser,  # noqa: F401
                                   SingleStepFewShotRecogniser)
from .feature_extractors import create_feature_extractor  # noqa: F401
from .classifier_heads import LinearClassifier, MeanPooler, PrototypicalClassifier  # noqa: F401
from .ops_counter import OpsCounter  # noqa: F401
from .evaluation import TestEvaluator, TrainEvaluator, ValidationEvaluator  # noqa: F401
from .data_utils import attach_frame_history, get_batch_indices, unpack_task  # noqa: F401

__all__ = ['OrbitError', 'load_library', 'FewShotRecogniser', 'MultiStepFewShotRecogniser',
           'SingleStepFewShotRecogniser', 'create_feature_extractor', 'LinearClassifier', 'MeanPooler',
           'PrototypicalClassifier', 'OpsCounter', 'TestEvaluator', 'TrainEvaluator', 'ValidationEvaluator',
           'attach_frame_history', 'get_batch_indices', 'unpack_task']
