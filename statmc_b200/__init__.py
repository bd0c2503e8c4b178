"""statmc_b200 -- B200-native (sm_100a) implementation of StatMC's data-parallel hot path:
per-pixel streaming moment accumulation + the statistical denoiser, behind a C ABI (include/statmc_b200.h).

Importing the package loads statmc_b200/libstatmc_b200.so (built in-tree by `python -m statmc_b200.build`);
there is no CPU or PyTorch fallback.
"""
from . import _capi  # noqa: F401  (fails loudly if the library is missing)
from .api import Buffer, Context, Denoiser, MomentState, PinnedArray, denoise_host  # noqa: F401

__all__ = ["Buffer", "Context", "Denoiser", "MomentState", "PinnedArray", "denoise_host"]
