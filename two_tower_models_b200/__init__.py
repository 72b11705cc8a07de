"""two_tower_models_b200: B200-native (sm_100a) drop-in for the hot path of gauravchak/two_tower_models.

Same class names, constructor arguments, methods and state_dict keys as the reference's
`src/` modules; the arithmetic runs in hand-written CUDA kernels behind the C ABI in
include/tt_b200.h (libtt_b200.so).  CUDA only - there is no CPU fallback.
"""
from . import _native, ops  # noqa: F401
from .history import UserHistoryEncoder  # noqa: F401
from .mips import BaselineMIPSModule  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .towers import TwoTowerBaseRetrieval, TwoTowerWithUserHistoryEncoder  # noqa: F401
from .debias import (  # noqa: F401
    TwoTowerWithDebiasing,
    TwoTowerWithPositionDebiasedWeights,
    TwoTowerWithUserDebiasedWeights,
)

__all__ = [
    "ops",
    "BaselineMIPSModule",
    "FusedAdam",
    "TwoTowerBaseRetrieval",
    "TwoTowerWithUserHistoryEncoder",
    "TwoTowerWithPositionDebiasedWeights",
    "TwoTowerWithUserDebiasedWeights",
    "TwoTowerWithDebiasing",
    "UserHistoryEncoder",
]
