"""spral_b200: B200-native numeric factorisation / solve engine for SPRAL SSIDS.

Only the hot path lives here: the CUDA engine behind the subtree plug-in C ABI
(csrc/, include/spral_ssids_b200.h) and the host-side mirror of the reference
interfaces (ssids.py).  There is no CPU implementation in this package.
"""
from . import _lib  # noqa: F401
from .ssids import (Analysis, SymbolicSubtree, NumericSubtree, analyse, factor, solve)  # noqa: F401
from ._lib import Options, Stats, Contrib  # noqa: F401
