"""Measurement operators mirroring ``pxmcmc/measurements.py`` of the reference.

numpy vector in -> numpy out; CUDA tensor ([n] or [nchains, n]) in -> tensor out."""
from warnings import warn

import numpy as np
import torch
from scipy import sparse

from . import device as D


class Measurement:
    """Base class (pxmcmc/measurements.py:7-35)."""

    def __init__(self, ndata, npix):
        self.ndata = ndata
        self.npix = npix

    def forward(self, X):
        raise NotImplementedError

    def adjoint(self, Y):
        raise NotImplementedError


def _veclen(X):
    return X.shape[-1] if D.is_dev(X) else len(X)


class Identity(Measurement):
    """What goes in comes out (pxmcmc/measurements.py:38-56; a rectangular
    ``eye(ndata, npix)`` truncates or zero-pads)."""

    _pxm_native = True

    def __init__(self, ndata, npix):
        super().__init__(ndata, npix)
        self.eye = sparse.eye(self.ndata, self.npix)
        self.eye_adj = self.eye.getH()

    @staticmethod
    def _resize(X, n_out):
        n_in = _veclen(X)
        if n_in == n_out:
            return X
        if D.is_dev(X):
            out = torch.zeros(X.shape[:-1] + (n_out,), dtype=X.dtype, device=X.device)
            k = min(n_in, n_out)
            out[..., :k] = X[..., :k]
            return out
        out = np.zeros(n_out, dtype=np.result_type(X, float))
        k = min(n_in, n_out)
        out[:k] = X[:k]
        return out

    def forward(self, X):
        assert _veclen(X) == self.npix
        return self._resize(X, self.ndata)

    def adjoint(self, Y):
        assert _veclen(Y) == self.ndata
        return self._resize(Y, self.npix)


class PathIntegral(Measurement):
    """y = A x and x~ = A^H y for a sparse path matrix (pxmcmc/measurements.py:59-83).
    Warp-per-row CSR SpMV on the device; like the reference, the (Hermitian)
    transpose is materialised once as its own CSR."""

    _pxm_native = True

    def __init__(self, path_matrix):
        self.path_matrix = path_matrix
        self.path_matrix_adj = self.path_matrix.getH()
        self.ndata, self.npix = path_matrix.shape
        self._dev = None

    def _upload(self):
        if self._dev is None:
            if np.iscomplexobj(self.path_matrix.data if sparse.issparse(self.path_matrix) else self.path_matrix):
                raise NotImplementedError("complex-valued path matrices are not supported on the device path")
            out = []
            for M in (sparse.csr_matrix(self.path_matrix), sparse.csr_matrix(self.path_matrix_adj)):
                M.sort_indices()
                dv = D.dev()
                out.append((torch.from_numpy(M.indptr.astype(np.int32)).to(dv),
                            torch.from_numpy(M.indices.astype(np.int32)).to(dv),
                            torch.from_numpy(M.data.astype(np.float64)).to(dv), M.shape[0], M.shape[1]))
            self._dev = out
        return self._dev

    def forward(self, X):
        assert _veclen(X) == self.npix
        ip, ix, v, nr, nc = self._upload()[0]
        return D.like_input(D.csr_spmv_dev(ip, ix, v, D.to_dev_c(X), nr, nc), X)

    def adjoint(self, Y):
        assert _veclen(Y) == self.ndata
        ip, ix, v, nr, nc = self._upload()[1]
        return D.like_input(D.csr_spmv_dev(ip, ix, v, D.to_dev_c(Y), nr, nc), Y)


class WeakLensingHarmonic(Measurement):
    """Spherical Kaiser-Squires kernel in harmonic space (pxmcmc/measurements.py:86-182)."""

    def __init__(self, L, mask=None, ngal=None):
        if L < 1:
            raise ValueError("Bandlimit {} must be greater than 0.".format(L))
        if L > 1024:
            warn("Bandlimit {} is very large, computational price is large.".format(L))
        self.L = L
        self.shape = (self.L ** 2,)
        self.harmonic_kernel = self.compute_harmonic_kernel()
        self.var_e = 0.37 ** 2

    def compute_harmonic_kernel(self):
        """k[l^2+l+m] = -sqrt((l+2)(l-1)/((l+1)l)) for l >= 2, ones below (zeroed on use)."""
        k = np.ones(self.L ** 2, dtype=float)
        for el in range(2, self.L):
            k[el * el: (el + 1) ** 2] = -1.0 * np.sqrt(((el + 2.0) * (el - 1.0)) / ((el + 1.0) * el))
        return k

    def _kernel_per_l(self):
        """per-degree multiplier with the l < 2 entries already zero (what harmonic_mapping applies)"""
        els = np.arange(self.L, dtype=float)
        g = np.zeros(self.L)
        g[2:] = -np.sqrt(((els[2:] + 2.0) * (els[2:] - 1.0)) / ((els[2:] + 1.0) * els[2:]))
        return g

    def harmonic_mapping(self, flm):
        out = flm * (self.harmonic_kernel if not D.is_dev(flm) else D.to_dev_f(self.harmonic_kernel))
        out[..., :4] = 0
        return out

    def harmonic_inverse_mapping(self, flm):
        out = flm / (self.harmonic_kernel if not D.is_dev(flm) else D.to_dev_f(self.harmonic_kernel))
        out[..., :4] = 0
        return out

    def forward(self, klm):
        return self.harmonic_mapping(klm)

    def adjoint(self, glm):
        return self.harmonic_mapping(glm)

    def sks_estimate(self, glm):
        return self.harmonic_inverse_mapping(glm)


class WeakLensing(WeakLensingHarmonic):
    """Pixel-space weak lensing operator (pxmcmc/measurements.py:185-304):
    kappa --spin-0 forward SHT--> klm --kernel--> glm --spin-2 inverse SHT--> gamma
    --mask gather--> --covariance weight-->, and the exact adjoint chain."""

    _pxm_native = True

    def __init__(self, L, mask=None, ngal=None):
        super().__init__(L, mask, ngal)
        self.shape = (self.L, 2 * self.L - 1)
        if mask is None:
            self.mask = np.ones(self.shape, dtype=bool)
        else:
            self.mask = np.asarray(mask).astype(bool)
        if self.mask.shape != self.shape:
            raise ValueError("Shape of mask map is incorrect!")
        if ngal is None:
            self.inv_cov = self.mask_forward(np.ones(self.shape))
        else:
            self.inv_cov = self.ngal_to_inv_cov(ngal)
        self.npix = self.L * (2 * self.L - 1)
        self.ndata = int(self.mask.sum())
        self._dev = None

    # -- host-side helpers kept for API compatibility -------------------------
    def mask_forward(self, f):
        if f is not f:
            raise ValueError("Signal is NaN.")
        if f.shape != self.shape:
            raise ValueError("Signal shape is incorrect for mw-sampling")
        return f[self.mask]

    def mask_adjoint(self, x):
        if x is not x:
            raise ValueError("Signal is NaN.")
        f = np.zeros(self.shape, dtype=complex)
        f[self.mask] = x
        return f

    def ngal_to_inv_cov(self, ngal):
        return np.sqrt((2.0 * self.mask_forward(np.asarray(ngal))) / (self.var_e))

    def cov_weight(self, x):
        return x * self.inv_cov

    # -- device path ------------------------------------------------------------
    def _upload(self):
        if self._dev is None:
            dv = D.dev()
            idx = torch.from_numpy(np.flatnonzero(self.mask.ravel()).astype(np.int32)).to(dv)
            self._dev = (idx, D.to_dev_f(self.inv_cov), D.to_dev_f(self._kernel_per_l()))
        return self._dev

    def _forward(self, kappa, masking=False, cov_weighting=False):
        idx, w, gl = self._upload()
        x = D.to_dev_c(kappa)
        if x.shape[-1] != self.npix:  # a map given as an (L, 2L-1) array, as the reference's np.reshape accepts
            x = x.reshape(-1) if x.numel() == self.npix else x.reshape(-1, self.npix)
        nb = 1 if x.dim() == 1 else x.shape[0]  # a [nchains, npix] batch stays a batch (also for one chain)
        klm = D.ShtPlan.get(self.L, 0, nb).forward(x)
        gamma = D.ShtPlan.get(self.L, 2, nb).inverse(klm, gl=gl)
        if masking:
            gamma = D.gather_dev(gamma, idx, w if cov_weighting else None, self.ndata)
        elif cov_weighting:
            gamma = gamma * w
        return D.like_input(gamma, kappa)

    def _adjoint(self, gamma, masking=False, cov_weighting=False):
        idx, w, gl = self._upload()
        y = D.to_dev_c(gamma)
        if masking:
            g = D.scatter_dev(y, idx, w if cov_weighting else None, self.npix)
        else:
            g = y
            if g.shape[-1] != self.npix:
                g = y.reshape(-1) if y.numel() == self.npix else y.reshape(-1, self.npix)
            if cov_weighting:
                g = g * w
        nb = 1 if g.dim() == 1 else g.shape[0]
        glm = D.ShtPlan.get(self.L, 2, nb).inverse_adjoint(g, gl=gl)
        kappa = D.ShtPlan.get(self.L, 0, nb).forward_adjoint(glm)
        return D.like_input(kappa, gamma)

    def forward(self, kappa):
        return self._forward(kappa, masking=True, cov_weighting=True)

    def adjoint(self, gamma):
        return self._adjoint(gamma, masking=True, cov_weighting=True)

    # the same operator with the convergence given / returned as harmonic coefficients: `forward` without its
    # leading spin-0 forward SHT (measurements.py:223), `adjoint` without its trailing forward_adjoint (:239)
    _pxm_harmonic_input = True

    def _forward_from_harmonic(self, klm):
        idx, w, gl = self._upload()
        x = D.to_dev_c(klm)
        nb = 1 if x.dim() == 1 else x.shape[0]
        gamma = D.ShtPlan.get(self.L, 2, nb).inverse(x, gl=gl)
        return D.like_input(D.gather_dev(gamma, idx, w, self.ndata), klm)

    def _adjoint_to_harmonic(self, gamma):
        idx, w, gl = self._upload()
        y = D.to_dev_c(gamma)
        nb = 1 if y.dim() == 1 else y.shape[0]
        g = D.scatter_dev(y, idx, w, self.npix)
        return D.like_input(D.ShtPlan.get(self.L, 2, nb).inverse_adjoint(g, gl=gl), gamma)


# the harmonic-space composition in ForwardOperator bypasses `forward` / `adjoint`: it is only taken while a
# (sub)class still uses these very implementations
WeakLensing._pxm_fused_methods = (WeakLensing.forward, WeakLensing.adjoint)
