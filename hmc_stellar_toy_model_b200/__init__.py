"""B200-native RHMC leapfrog hot path of the stellar toy model, behind the reference's Python call surface.

    from hmc_stellar_toy_model_b200.sampler_RHMC import multi_gym, single_gym     # drop-in classes
    from hmc_stellar_toy_model_b200.context import RHMCContext                    # batch API (many chains/fields)

Everything numeric on the path runs in hand-written sm_100a CUDA kernels reached through the C ABI of
libstellar_rhmc.so (include/stellar_rhmc.h).  There is no CPU fallback.
"""
from . import _capi  # noqa: F401
from .context import RHMCContext, RunResult  # noqa: F401

__all__ = ["RHMCContext", "RunResult"]
