"""Transforms mirroring ``pxmcmc/transforms.py`` of the reference.

Every method accepts a 1-D numpy vector (numpy out, as in the reference) or a
CUDA tensor of shape [n] / [nchains, n] (tensor out, stays in HBM; this is what
the samplers use)."""
import numpy as np

from . import device as D
from .utils import j_max


class Transform:
    """Base class (pxmcmc/transforms.py:8-33)."""

    def forward(self, X):
        raise NotImplementedError

    def inverse(self, X):
        raise NotImplementedError

    def forward_adjoint(self, X):
        raise NotImplementedError

    def inverse_adjoint(self, X):
        raise NotImplementedError


class IdentityTransform(Transform):
    """Does nothing (pxmcmc/transforms.py:36-56)."""

    _pxm_native = True

    def __init__(self):
        pass

    def forward(self, X):
        return X

    def forward_adjoint(self, X):
        return X

    def inverse(self, X):
        return X

    def inverse_adjoint(self, X):
        return X


class SphericalWaveletTransform(Transform):
    """Axisymmetric scale-discretised wavelet transform on MW sampling
    (pxmcmc/transforms.py:59-166): multiresolution, coefficient vector =
    [scaling map, wavelet maps j = J_min..J_max].

    ``forward`` = pys2let.analysis_px2wav, ``inverse`` = synthesis_wav2px,
    ``inverse_adjoint`` = synthesis_adjoint_px2wav, ``forward_adjoint`` =
    analysis_adjoint_wav2px, each one launch sequence of libpxmcmc_b200
    (ring FFT -> DMMA Legendre over rings -> DMMA Legendre over l -> ring FFT)
    covering all scales at once.
    """

    _pxm_native = True

    def __init__(self, L, B, J_min, dirs=1, spin=0, harmonic=False, nchains=1):
        if dirs != 1 or spin != 0:
            raise NotImplementedError("only dirs=1, spin=0 (what the reference's drivers use) is implemented")
        if harmonic:
            # the reference's harmonic=True branch points at pys2let functions that do not exist
            raise NotImplementedError("harmonic=True is not available (dead branch in the reference)")
        self.L = L
        self.B = B
        self.J_min = J_min
        self.J_max = j_max(L, B)
        self.nscales = self.J_max - self.J_min + 1
        self.dirs = dirs
        self.spin = spin
        self.params = {"B": B, "L": L, "J_min": J_min, "N": dirs, "spin": spin, "upsample": 0}
        self._plans = {}
        self._get_ncoefs(nchains)

    def _plan(self, nb):
        if nb not in self._plans:
            self._plans[nb] = D.WaveletPlan.get(self.L, self.B, self.J_min, nb)
        return self._plans[nb]

    def _get_ncoefs(self, nchains=1):
        p = self._plan(nchains)
        self.bandlimits = list(p.bandlimits)
        self.nscal = p.nscal
        self.nwav = p.ncoefs - p.nscal
        self.ncoefs = p.ncoefs

    def _apply(self, name, X):
        x = D.to_dev_c(X)
        nb = 1 if x.dim() == 1 else x.shape[0]
        return D.like_input(getattr(self._plan(nb), name)(x), X)

    # harmonic-space ends of the synthesis pair (used by ForwardOperator when the measurement starts with a
    # spin-0 forward SHT: A_fwd(L,0) o A_inv(L,0) = I on f_lm)
    def _inverse_harmonic(self, X):
        """wavelet coefficients -> f_lm of the image (``inverse`` without its final inverse SHT)"""
        return self._apply("synthesis_harmonic", X)

    def _inverse_adjoint_harmonic(self, flm):
        """adjoint of ``_inverse_harmonic``"""
        return self._apply("synthesis_adjoint_harmonic", flm)

    def forward(self, X):
        """image -> wavelet coefficients"""
        return self._apply("analysis", X)

    def inverse(self, X):
        """wavelet coefficients -> image"""
        return self._apply("synthesis", X)

    def inverse_adjoint(self, X):
        """image -> wavelet coefficients (adjoint of ``inverse``)"""
        return self._apply("synthesis_adjoint", X)

    def forward_adjoint(self, X):
        """wavelet coefficients -> image (adjoint of ``forward``)"""
        return self._apply("analysis_adjoint", X)
