"""pxmcmc_b200 -- B200-native drop-in for the per-iteration proximal-Langevin
path of auggiemarignier/pxmcmc.

    import pxmcmc_b200 as pxmcmc
    from pxmcmc_b200.mcmc import MYULA, PxMALA, SKROCK, PxMCMCParams
    from pxmcmc_b200.forward import SphericalWaveletTransformOperator

Same class / method / attribute names as the reference's ``pxmcmc`` package for
the hot path (mcmc, forward, transforms, measurements, prior, utils); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared
in ``include/pxmcmc_b200.h``.  There is no CPU fallback: the first call that
needs the device raises if the shared library or a Blackwell GPU is missing.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (loads libpxmcmc_b200.so; raises loudly if it is not built)
