"""elisa_b200 -- B200-native (sm_100a) batched forward-folding likelihood + gradient for
wcxve/elisa's data-parallel hot path.  See DESIGN.md / INTEGRATION.md.

The compute lives in ``lib/libelisa_b200.so`` (hand-written CUDA behind the C ABI of
``include/elisa_b200.h``); this package is the thin Python host side that mirrors the
reference's model / data / likelihood interfaces for that path.  No CPU fallback."""

from . import models
from .data import FixedData
from .likelihood import Likelihood, sliced_gemm
from .models import *  # noqa: F401,F403
from ._lib import ElisaB200Error, load as load_library

__version__ = '0.1.0'
