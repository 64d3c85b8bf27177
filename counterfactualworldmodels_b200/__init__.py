"""counterfactualworldmodels_b200 -- the CWM masked video-autoencoder (VMAE) forward path, B200-native.

Only what the hot path needs: ``vmae`` (the drop-in predictor), ``conjoined_vmae`` (+ ``transformer``,
``preprocessor``: the padded / IMU-conditioned conjoined predictors of BASELINE config 5), ``prediction`` (the two
wrapper steps that bracket the predictor), ``dist`` (sharding of counterfactual batches over the GPUs of one box) and
``csrc`` (hand-written sm_100a CUDA behind the C ABI of ``include/cwm_b200.h``).
"""
from . import _lib  # noqa: F401
from .vmae import (PretrainVisionTransformer, base_4x4patch_2frames_1tube, base_8x8patch_2frames_1tube,  # noqa: F401
                   base_16x16patch_2frames_1tube, large_4x4patch_2frames_1tube, compact_mask,
                   get_sinusoid_encoding_table)
from .conjoined_vmae import (ConjoinedPaddedVisionTransformer, ConjoinedPretrainVisionTransformer, ImuEncoder,  # noqa: F401
                             PaddedVisionTransformer, imu400_8x8patch_2frames_1tube_flowbackrgb01,
                             imu400_base_4x4patch_2frames_1tube)
from .prediction import PredictorBasedGenerator, RectangularizeMasks, unpatchify_scatter  # noqa: F401

__version__ = "0.1.0"
