"""thejoker_b200: a B200-native (sm_100a, FP64) implementation of The Joker's hot
path behind the reference's Python API.

    prior cache -> Kepler solve per epoch -> design matrix -> Gaussian-marginal
    log-likelihood -> rejection accept -> linear-parameter draw

The compute lives in libthejoker_b200.so (hand-written CUDA, C ABI in
include/thejoker_b200.h) reached through ctypes; torch is used for device buffers,
streams and torch.distributed only.  There is no CPU fallback.
"""
from . import units  # noqa: F401
from .data import RVData  # noqa: F401
from .helper import CJokerHelper, extract_spec  # noqa: F401
from .multistar import MultiStarJoker  # noqa: F401
from .prior import JokerPrior  # noqa: F401
from .samples import JokerSamples  # noqa: F401
from .thejoker import TheJoker  # noqa: F401

__version__ = "0.1.0"
__all__ = ["TheJoker", "RVData", "JokerPrior", "JokerSamples", "CJokerHelper", "MultiStarJoker",
           "units"]
