"""titanet_b200 -- the TitaNet speaker-embedding hot path on NVIDIA B200 (sm_100a).

Same module / class names as the reference project's ``src/`` (``modules``, ``models``,
``losses``, ``transforms``), hand-written CUDA kernels underneath
(``libtitanet_sm100.so``, C ABI in ``include/titanet_b200.h``).  No CPU fallback.
"""
from . import losses, models, modules, transforms  # noqa: F401
from ._lib import LIB_PATH, TitanetLibraryError  # noqa: F401
from .models import TitaNet  # noqa: F401

__version__ = "0.1.0"
