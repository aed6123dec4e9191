"""CPU restatement of the reference's Python layer for the proximal-Langevin
step.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it follows.  PINNED: the committed
fixtures ``tests/golden/ref_*.npz`` were produced by the unmodified reference
(``oracle/gen_golden.py``) and ``tests/test_oracle_golden.py`` checks this file
against them.  The reference tree does not exist on the GPU box, which is why
this restatement (and not the reference itself) is what the ``-m gpu`` tests
compare against.
"""
from math import pi, sqrt

import numpy as np

from . import s2let_ref, ssht_ref


# ---------------------------------------------------------------- utils.py
def soft(X, T):
    """pxmcmc/utils.py:55-67 + _sign :84-88.  sign(x)(|x|-T), 0 where |x| <= T."""
    X = np.array(X)
    a = np.abs(X)
    safe = np.where(a == 0, 1.0, a)
    out = (np.where(a == 0, 0, X) / safe) * (a - T)
    out[a <= T] = 0
    return out


def chebyshev1(x, k):
    """pxmcmc/utils.py:128-151."""
    t0, t1 = 1, x
    if k == 0:
        return t0
    for _ in range(k - 1):
        t0, t1 = t1, 2 * x * t1 - t0
    return t1


def chebyshev2(x, k):
    """pxmcmc/utils.py:154-177."""
    u0, u1 = 1, 2 * x
    if k == 0:
        return u0
    for _ in range(k - 1):
        u0, u1 = u1, 2 * x * u1 - u0
    return u1


def cheb1der(x, k):
    """pxmcmc/utils.py:180-197."""
    return 0 if k == 0 else k * chebyshev2(x, k - 1)


def mw_map_weights(L):
    """pxmcmc/utils.py:249-283: exact MW quadrature weights incl. 2 pi/(2L-1)."""
    n = 2 * L - 1
    w = np.array([ssht_ref.mw_quad_weight(m) * np.exp(-1j * m * pi / n) for m in range(-(L - 1), L)])
    wr = (np.fft.fft(np.fft.ifftshift(w)) * 2 * pi / n ** 2).real
    q = wr[:L].copy()
    q[: L - 1] += wr[n - 1: L - 1: -1]
    return np.repeat(q, n)


# ---------------------------------------------------------------- transforms.py
class WaveletTransform:
    """pxmcmc/transforms.py:59-166 (N=1, spin=0, upsample=0).  Coefficient vector =
    [scaling map, wavelet maps by increasing j] (utils.flatten_mlm :11-22)."""

    def __init__(self, L, B, J_min):
        self.L, self.B, self.J_min = L, B, J_min
        self.bandlimits = s2let_ref.bandlimits(B, L, J_min)
        self.nscal = s2let_ref.mw_size(self.bandlimits[0])
        self.nwav = sum(s2let_ref.mw_size(b) for b in self.bandlimits[1:])
        self.ncoefs = self.nscal + self.nwav

    def forward(self, X):  # :102-112
        w, s = s2let_ref.analysis_px2wav(np.asarray(X).astype(complex), self.B, self.L, self.J_min)
        return np.concatenate((s, w))

    def inverse(self, X):  # :114-127
        X = np.asarray(X).astype(complex)
        return s2let_ref.synthesis_wav2px(X[self.nscal:], X[: self.nscal], self.B, self.L, self.J_min)

    def inverse_adjoint(self, X):  # :129-139
        w, s = s2let_ref.synthesis_adjoint_px2wav(np.asarray(X).astype(complex), self.B, self.L, self.J_min)
        return np.concatenate((s, w))

    def forward_adjoint(self, X):  # :141-154
        X = np.asarray(X).astype(complex)
        return s2let_ref.analysis_adjoint_wav2px(X[self.nscal:], X[: self.nscal], self.B, self.L, self.J_min)


class IdentityTransform:
    """pxmcmc/transforms.py:36-56."""

    def forward(self, X):
        return X

    inverse = forward_adjoint = inverse_adjoint = forward


# ---------------------------------------------------------------- measurements.py
class IdentityMeasurement:
    """pxmcmc/measurements.py:38-56 (sparse eye(ndata, npix) and its transpose)."""

    def __init__(self, ndata, npix):
        self.ndata, self.npix = ndata, npix

    def forward(self, X):
        assert len(X) == self.npix
        out = np.zeros(self.ndata, dtype=np.result_type(X, float))
        k = min(self.ndata, self.npix)
        out[:k] = X[:k]
        return out

    def adjoint(self, Y):
        assert len(Y) == self.ndata
        out = np.zeros(self.npix, dtype=np.result_type(Y, float))
        k = min(self.ndata, self.npix)
        out[:k] = Y[:k]
        return out


class PathIntegral:
    """pxmcmc/measurements.py:59-83: y = A x, x~ = A^H y (A scipy sparse)."""

    def __init__(self, path_matrix):
        self.A = path_matrix.tocsr()
        self.AH = self.A.conj().T.tocsr()
        self.ndata, self.npix = self.A.shape

    def forward(self, X):
        assert len(X) == self.npix
        return self.A.dot(X)

    def adjoint(self, Y):
        assert len(Y) == self.ndata
        return self.AH.dot(Y)


def wl_harmonic_kernel(L):
    """pxmcmc/measurements.py:151-171: -sqrt((l+2)(l-1)/((l+1)l)) for l >= 2, first
    four entries (l = 0, 1) forced to zero by harmonic_mapping."""
    k = np.zeros(L * L)
    for l in range(2, L):
        k[l * l: (l + 1) * (l + 1)] = -sqrt(((l + 2.0) * (l - 1.0)) / ((l + 1.0) * l))
    return k


class WeakLensing:
    """pxmcmc/measurements.py:185-304."""

    def __init__(self, L, mask=None, ngal=None):
        self.L = L
        self.shape = (L, 2 * L - 1)
        self.kernel = wl_harmonic_kernel(L)
        self.mask = np.ones(self.shape, dtype=bool) if mask is None else np.asarray(mask).astype(bool)
        if self.mask.shape != self.shape:
            raise ValueError("Shape of mask map is incorrect!")
        if ngal is None:
            self.inv_cov = np.ones(self.shape)[self.mask]  # :202
        else:
            self.inv_cov = np.sqrt(2.0 * np.asarray(ngal)[self.mask] / 0.37 ** 2)  # :282-293
        self.ndata, self.npix = int(self.mask.sum()), L * (2 * L - 1)

    def forward(self, kappa):  # :209-230
        klm = ssht_ref.forward(np.asarray(kappa).reshape(self.shape), self.L, 0)
        gamma = ssht_ref.inverse(klm * self.kernel, self.L, 2)
        return (gamma[self.mask] * self.inv_cov).ravel()

    def adjoint(self, gamma):  # :218-240
        g = np.zeros(self.shape, dtype=complex)
        g[self.mask] = np.asarray(gamma) * self.inv_cov
        glm = ssht_ref.inverse_adjoint(g, self.L, 2)
        return ssht_ref.forward_adjoint(glm * self.kernel, self.L, 0).ravel()


# ---------------------------------------------------------------- forward.py
def inverse_covariance(data, sig_d):
    """pxmcmc/forward.py:74-88 for scalar / vector sigma: returns the DIAGONAL of
    the inverse covariance.  Real sigma with complex data => var*(1+i)/sqrt(2)."""
    var = np.asarray(sig_d, dtype=float) ** 2 if not np.iscomplexobj(sig_d) else np.asarray(sig_d) ** 2
    if np.iscomplexobj(data) and not np.iscomplexobj(var):
        var = var / np.sqrt(2) * (1 + 1j)
    if np.ndim(var) == 0:
        return np.full(len(data), 1 / var)
    if var.ndim == 1 and var.size == len(data):
        return 1 / var
    raise TypeError("sig_d must be a float scalar, vector or 2D matrix")


class ForwardOperator:
    """pxmcmc/forward.py:9-88 (diagonal covariance only)."""

    def __init__(self, data, sig_d, setting, transform, measurement, nparams):
        if setting not in ("analysis", "synthesis"):
            raise ValueError
        self.data = np.asarray(data)
        self.invcov = inverse_covariance(self.data, sig_d)
        self.setting, self.transform, self.measurement, self.nparams = setting, transform, measurement, nparams

    def forward(self, X):  # :36-46, :60-64
        if self.setting == "analysis":
            return self.measurement.forward(X)
        return self.measurement.forward(self.transform.inverse(X))

    def calc_gradg(self, preds):  # :48-58, :66-72
        g = self.measurement.adjoint(self.invcov * (preds - self.data))
        if self.setting == "synthesis":
            g = self.transform.inverse_adjoint(g)
        return g


# ---------------------------------------------------------------- prior.py
class L1:
    """pxmcmc/prior.py:8-53."""

    def __init__(self, setting, fwd, adj, T):
        assert setting in ("analysis", "synthesis")
        self.setting, self.fwd, self.adj, self.T = setting, fwd, adj, T

    def prior(self, X):  # :28-35
        return np.sum(np.abs(X))

    def proxf(self, X):  # :37-53
        if self.setting == "synthesis":
            return soft(X, self.T)
        a = self.adj(X)
        return X + self.fwd(soft(a, self.T) - a)


class S2WaveletsL1(L1):
    """pxmcmc/prior.py:56-84: threshold and prior weighted by the MW quadrature
    weights of every scale (synthesis only)."""

    def __init__(self, setting, fwd, adj, T, L, B, J_min):
        super().__init__(setting, fwd, adj, T)
        if setting != "synthesis":
            raise NotImplementedError
        self.map_weights = np.concatenate([mw_map_weights(b) for b in s2let_ref.bandlimits(B, L, J_min)])
        self.T = self.T * self.map_weights

    def prior(self, X):
        return np.sum(np.abs(self.map_weights * X))


class S2WaveletsL1PowerWeights(S2WaveletsL1):
    """pxmcmc/prior.py:87-149.  T is multiplied by BOTH weight sets (:81 then :108)
    and prior() applies the power weights twice (:110-111 -> :83-84)."""

    def __init__(self, setting, fwd, adj, T, L, B, J_min, eta=1):
        super().__init__(setting, fwd, adj, T, L, B, J_min)
        phi_l, psi_lm = s2let_ref.wavelet_tiling(B, L, 1, J_min, 0)
        bls = s2let_ref.bandlimits(B, L, J_min)
        ws = []
        # scaling part :120-128
        p = np.vdot(phi_l, phi_l).real
        Le = int(np.nonzero(phi_l)[0].max()) + 1
        th, _ = ssht_ref.sample_positions(Le)
        ws.append(np.repeat(2 * pi ** 2 / (p * ssht_ref.sample_length(Le)) * np.sin(th), 2 * Le - 1))
        # wavelet part :130-149
        ls = np.arange(L)
        for j, Le in enumerate(bls[1:]):
            col = psi_lm[:, j]
            p = np.vdot(col, col).real
            peak = int(np.argmax(col[ls * ls + ls]))
            th, _ = ssht_ref.sample_positions(Le)
            ws.append(np.repeat(2 * pi ** 2 * peak ** eta / (p * ssht_ref.sample_length(Le)) * np.sin(th), 2 * Le - 1))
        self.map_weights = np.concatenate(ws)
        self.T = self.T * self.map_weights

    def prior(self, X):
        return np.sum(np.abs(self.map_weights * (self.map_weights * X)))


# ---------------------------------------------------------------- mcmc.py
def logpi(fwd, prior, mu, X, preds):
    """pxmcmc/mcmc.py:71-82 (vdot conjugates its first argument; no 1/2)."""
    diff = fwd.data - preds
    L2 = np.vdot(diff, fwd.invcov * diff)
    pr = prior.prior(X)
    return -mu * pr - L2, L2, pr


def gradlogpi(fwd, prior, lmda, X):
    """pxmcmc/mcmc.py:84-89."""
    return -(X - prior.proxf(X)) / lmda - fwd.calc_gradg(fwd.forward(X))


def myula_step(X, proxf, gradg, delta, lmda, w):
    """pxmcmc/mcmc.py:185-201 with the Gaussian draw `w` injected."""
    return (1 - delta / lmda) * X + (delta / lmda) * proxf - delta * gradg + np.sqrt(2 * delta) * w


def myula_iteration(fwd, prior, delta, lmda, X, preds, w):
    """One pass of the loop body pxmcmc/mcmc.py:158-164."""
    gradg = fwd.calc_gradg(preds)
    px = prior.proxf(X)
    Xn = myula_step(X, px, gradg, delta, lmda, w)
    return Xn, fwd.forward(Xn)


def pxmala_logtransition(X1, X2, proxf, gradg, delta, lmda):
    """pxmcmc/mcmc.py:281-289, exactly as coded: (1/2*delta) == delta/2 and the
    (non-conjugated) sum of squares is squared again."""
    glp = -((X1 - proxf) / lmda) - gradg
    return -(1 / 2 * delta) * np.sum((X2 - X1 - (delta / 2) * glp) ** 2) ** 2


def pxmala_tune_delta(delta, lmda, accepted, i):
    """pxmcmc/mcmc.py:277-279."""
    d = delta * (1 + (accepted - 0.5) / ((i + 1) ** 0.75))
    return min(max(d, lmda * 1e-8), lmda / 2)


def skrock_coefs(s, eta=0.05):
    """pxmcmc/mcmc.py:300-306, :370-383 (as coded: ratio uses T_j(omega_1); k_j = 1)."""
    w0 = 1 + eta / (s * s)
    w1 = chebyshev1(w0, s) / cheb1der(w0, s)
    mus, nus, ks = np.zeros(s + 1), np.zeros(s + 1), np.zeros(s + 1)
    mus[1], nus[1], ks[1] = w1 / w0, s * w1 / 2, s * w1 / w0
    for j in range(2, s + 1):
        r = chebyshev1(w0, j - 1) / chebyshev1(w1, j)
        mus[j], nus[j], ks[j] = 2 * w1 * r, 2 * w0 * r, 1 - nus[0]
    return w0, w1, mus, nus, ks


def skrock_step(fwd, prior, delta, lmda, s, X, Z):
    """pxmcmc/mcmc.py:338-368 evaluated bottom-up (K_0..K_s): the reference's
    recursion is deterministic in (X, Z), so memoising gives identical values."""
    _, _, mus, nus, ks = skrock_coefs(s)
    sq = np.sqrt(2 * delta)
    K = [X]
    if s >= 1:
        K.append(X + mus[1] * delta * gradlogpi(fwd, prior, lmda, X + nus[1] * sq * Z) + ks[1] * sq * Z)
    for j in range(2, s + 1):
        K.append(mus[j] * delta * gradlogpi(fwd, prior, lmda, K[j - 1]) + nus[j] * K[j - 1] + ks[j] - K[j - 2])
    return K[s]
