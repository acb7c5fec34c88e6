"""B200-native (sm_100a) implementation of the per-clip audio-visual inference hot path of
parshva-mehta/multimodal-emotion-detection: log-mel/MFCC front-end -> encoders -> cross-modal attention fusion ->
T-pass MC-dropout uncertainty.  Python modules keep the reference's signatures (src/encoders.py,
src/attention.py, src/fusion.py, src/uncertainty.py, src/dataprocessing.py) and call hand-written CUDA kernels
through the C ABI in include/avhot.h (libavhot.so).  There is no CPU fallback."""
from . import _lib, _rng, dataprocessing, encoders, attention, fusion, uncertainty, synth, parallel, hotpath, data, exact  # noqa: F401
from ._rng import philox_context  # noqa: F401
from .encoders import SequenceEncoder, FrameEncoder, build_encoder  # noqa: F401
from .attention import CrossModalAttention, TemporalAttention, PairwiseModalityAttention  # noqa: F401
from .fusion import HybridFusion, EarlyFusion, LateFusion, LateFusionWithUncertainty, ConcatHead, build_fusion_model  # noqa: F401
from .uncertainty import (MCDropoutUncertainty, CalibrationMetrics, UncertaintyWeightedFusion,  # noqa: F401
                          compute_calibration_metrics, mc_reduce)  # noqa: F401
from .dataprocessing import condition_audio, logmel_mfcc, load_raw_audio  # noqa: F401
from .hotpath import AVHotPath  # noqa: F401
from .exact import NearTieResolver  # noqa: F401
from .data import SplitReader, save_splits_to_disk  # noqa: F401

__version__ = "0.1.0"
