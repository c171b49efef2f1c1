"""gdmix_b200 -- B200-native (sm_100a) random-effect / fixed-effect LR trainer behind GDMix's plugin API.

Importing the package loads lib/libgdmix_b200.so (hand-written CUDA, C ABI in include/gdmix_b200.h).
There is no CPU fallback: a missing library is an ImportError.
"""
from . import _capi  # noqa: F401  (loads the native library or raises)

__version__ = "0.1.0"

from .fixed_effect import FixedEffectLRLBFGSModel, FixedEffectLRModelLBFGS  # noqa: E402,F401
from .random_effect import RandomEffectLRLBFGSModel  # noqa: E402,F401
