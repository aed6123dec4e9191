"""Priors and proximal maps mirroring ``pxmcmc/prior.py`` of the reference."""
import numpy as np

from . import device as D
from .utils import _multires_bandlimits, j_max, mw_map_weights, mw_sample_positions, wavelet_tiling


class L1:
    """Laplace / L1 prior; its prox is soft thresholding (pxmcmc/prior.py:8-53)."""

    _pxm_native = True

    def __init__(self, setting, fwd, adj, T):
        assert setting in ["analysis", "synthesis"]
        self.setting = setting
        self.fwd = fwd
        self.adj = adj
        self.T = T
        self._Tdev = None

    # threshold on the device: (vector or None, scalar)
    def _T_args(self):
        if np.ndim(self.T) == 0:
            return None, float(self.T)
        if self._Tdev is None or self._Tdev[1] is not self.T:
            self._Tdev = (D.to_dev_f(np.asarray(self.T, dtype=float)), self.T)
        return self._Tdev[0], 0.0

    def _weights_dev(self):
        return None

    def prior(self, X):
        """sum |X| (pxmcmc/prior.py:28-35); one value per chain for a [nchains, n] tensor"""
        x = D.to_dev_c(X)
        r = D.reduce_dev(0, x, w=self._weights_dev()).real
        if D.is_dev(X):
            return r if x.dim() == 2 else r[0]
        return float(r[0].item())

    def proxf(self, X):
        if self.setting == "synthesis":
            return self._proxf_synthesis(X)
        return self._proxf_analysis(X)

    def _proxf_synthesis(self, X):
        Tv, Ts = self._T_args()
        x = D.to_dev_c(X) if (D.is_dev(X) or np.iscomplexobj(X)) else D.to_dev_f(X)
        return D.like_input(D.soft_dev(x, Tv, Ts), X)

    def _proxf_analysis(self, X):
        """X + fwd(soft(adj(X), T) - adj(X)) (pxmcmc/prior.py:52-53; adj(X) evaluated once)"""
        x = D.to_dev_c(X)
        a = D.to_dev_c(self.adj(x))
        Tv, Ts = self._T_args()
        s = D.soft_dev(a, Tv, Ts)
        diff = D.lincomb_dev([(1.0, s), (-1.0, a)])
        back = D.to_dev_c(self.fwd(diff))
        return D.like_input(D.lincomb_dev([(1.0, x), (1.0, back)]), X)


def is_library_l1(prior):
    """True when `prior` is an L1 (sub)class whose prox / prior are still the library's implementations.  The samplers
    take their fused device paths (soft threshold inside the update kernel, sum |w X| as a device reduction) only
    then; a user subclass that overrides `proxf`, `_proxf_synthesis`, `_proxf_analysis` or `prior` -- the reference's
    extension pattern -- is called through its own methods instead."""
    if not isinstance(prior, L1):
        return False
    t = type(prior)
    return (t.proxf is L1.proxf and t._proxf_synthesis is L1._proxf_synthesis and t._proxf_analysis is L1._proxf_analysis
            and t.prior in _LIBRARY_PRIOR_METHODS and t._weights_dev in _LIBRARY_WEIGHT_METHODS)


class S2_Wavelets_L1(L1):
    """L1 on spherical wavelet coefficients weighted by the exact MW quadrature
    weights of every scale (pxmcmc/prior.py:56-84)."""

    def __init__(self, setting, fwd, adj, T, L, B, J_min, dirs=1, spin=0):
        super().__init__(setting, fwd, adj, T)
        self.L = L
        self.B = B
        self.J_min = J_min
        self.J_max = j_max(L, B)
        self.nscales = self.J_max - J_min + 1
        self.dirs = dirs
        self.spin = spin
        if setting == "synthesis":
            bls = _multires_bandlimits(L, B, J_min, dirs, spin)
            self.map_weights = np.concatenate([mw_map_weights(el) for el in bls])
        else:
            raise NotImplementedError
        self.T = self.T * self.map_weights
        self._wdev = None

    def _weights_dev(self):
        if self._wdev is None or self._wdev[1] is not self.map_weights:
            self._wdev = (D.to_dev_f(self._prior_weights()), self.map_weights)
        return self._wdev[0]

    def _prior_weights(self):
        return self.map_weights

    def prior(self, X):
        """sum |w X| (pxmcmc/prior.py:83-84)"""
        return super().prior(X)


class S2_Wavelets_L1_Power_Weights(S2_Wavelets_L1):
    """Pixel-area x wavelet-power weighting, eqs 33-34 of Wallis et al. 2017
    (pxmcmc/prior.py:87-149).  As in the reference the threshold ends up multiplied
    by both weight sets and ``prior`` applies the power weights twice."""

    def __init__(self, setting, fwd, adj, T, L, B, J_min, dirs=1, spin=0, eta=1):
        super().__init__(setting, fwd, adj, T, L, B, J_min, dirs, spin)
        self.eta = eta
        if setting == "synthesis":
            self._get_weights()
        else:
            raise NotImplementedError
        self.T = self.T * self.map_weights
        self._wdev = None

    def _prior_weights(self):
        return self.map_weights * self.map_weights

    def _ring_weights(self, effective_L, scale):
        thetas, _ = mw_sample_positions(effective_L)
        return np.repeat(scale * np.sin(thetas), 2 * effective_L - 1)

    def _get_weights(self):
        phi_l, psi_lm = wavelet_tiling(self.B, self.L, self.dirs, self.J_min, self.spin)
        parts = [self._calculate_scaling_weights(phi_l)]
        parts += self._calculate_wavelet_weights(psi_lm)
        self.map_weights = np.concatenate(parts)

    def _calculate_scaling_weights(self, phi_l):
        power = np.vdot(phi_l, phi_l).real
        effective_L = int(np.nonzero(phi_l)[0].max()) + 1
        nsamples = effective_L * (2 * effective_L - 1)
        return self._ring_weights(effective_L, 2 * np.pi ** 2 / (power * nsamples))

    def _calculate_wavelet_weights(self, psi_lm):
        bls = _multires_bandlimits(self.L, self.B, self.J_min)
        ls = np.arange(self.L)
        out = []
        for j, effective_L in enumerate(bls[1:]):
            col = psi_lm[:, j]
            power = np.vdot(col, col).real
            peak_l = int(np.argmax(col[ls * ls + ls]))
            nsamples = int(effective_L) * (2 * int(effective_L) - 1)
            out.append(self._ring_weights(int(effective_L), (2 * np.pi ** 2) * (peak_l ** self.eta) / (power * nsamples)))
        return out


_LIBRARY_PRIOR_METHODS = (L1.prior, S2_Wavelets_L1.prior)
_LIBRARY_WEIGHT_METHODS = (L1._weights_dev, S2_Wavelets_L1._weights_dev)
