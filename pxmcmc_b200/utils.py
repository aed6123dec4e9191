"""Helpers mirroring ``pxmcmc/utils.py`` of the reference.

Array-sized operations (``soft``) run on the GPU through the C ABI; the small
setup-time formulas (quadrature weights, Chebyshev polynomials, pixel areas)
stay on the host exactly as in the reference, where they are plain numpy too.
"""
import numpy as np

from . import _lib


# ------------------------------------------------------------------ layout
def flatten_mlm(wav_lm, scal_lm):
    """Scaling coefficients first, then the wavelet coefficients scale by scale
    (reference ``utils.flatten_mlm``, pxmcmc/utils.py:11-22; a 2-D ``wav_lm`` is
    unrolled column by column)."""
    from .device import is_dev

    if is_dev(wav_lm):
        import torch

        w = wav_lm.t().reshape(-1) if wav_lm.dim() == 2 else wav_lm.reshape(-1)
        return torch.cat((scal_lm, w))
    w = np.asarray(wav_lm)
    return np.concatenate((np.asarray(scal_lm), w.ravel(order="F")))


def expand_mlm(mlm, nscales=None, nscalcoefs=None, flatten_wavs=False):
    """Inverse of :func:`flatten_mlm` (pxmcmc/utils.py:25-52): returns
    ``(wavelet coefficients, scaling coefficients)``."""
    if (nscales is None) == (nscalcoefs is None):
        if nscales is None:
            raise ValueError("Set either 'nscales', or 'nscalcoefs'")
        raise ValueError("Give only one of 'nscales' or 'nscalcoefs'")
    if nscalcoefs is not None:
        return mlm[nscalcoefs:], mlm[:nscalcoefs]
    width = mlm.size // (nscales + 1)
    assert width > 0
    scal = mlm[:width]
    wav = np.zeros((width, nscales), dtype=complex)
    for k in range(nscales):
        wav[:, k] = mlm[(k + 1) * width: (k + 2) * width]
    if flatten_wavs:
        wav = wav.ravel(order="F")
    return wav, scal


# ------------------------------------------------------------------ thresholding
def soft(X, T=0.1):
    """Soft thresholding sign(x)(|x|-T), zero where |x| <= T (pxmcmc/utils.py:55-67).
    numpy / list in -> numpy out (computed on the GPU); device tensor in -> device tensor out."""
    from . import device as D

    if D.is_dev(X):
        Tv, Ts = D._T_args(T)
        return D.soft_dev(X.contiguous(), Tv, Ts)
    a = np.array(X)
    shape = a.shape
    if np.iscomplexobj(a):
        x = D.to_dev_c(a.ravel())
    else:
        x = D.to_dev_f(a.ravel())
    if np.ndim(T) == 0:
        Tv, Ts = None, float(T)
    else:
        Tv, Ts = D.to_dev_f(np.broadcast_to(np.asarray(T, dtype=float), shape).ravel()), 0.0
    return D.to_host(D.soft_dev(x, Tv, Ts)).reshape(shape)


def hard(X, T=0.1):
    """Keep the largest 100*T % of |X| (pxmcmc/utils.py:70-81; modifies X in place
    like the reference).  Not on the sampler's path; host numpy."""
    mags = np.sort(np.abs(X))
    cut = mags[-int(T * len(X))]
    X[np.abs(X) < cut] = 0
    return X


# ------------------------------------------------------------------ Chebyshev
def chebyshev1(X, order):
    """T_order(X) by the three-term recurrence (pxmcmc/utils.py:128-151)."""
    if order < 0:
        raise ValueError("order must be >= 0")
    prev, cur = 1, X
    if order == 0:
        return prev
    for _ in range(order - 1):
        prev, cur = cur, 2 * X * cur - prev
    return cur


def chebyshev2(X, order):
    """U_order(X) (pxmcmc/utils.py:154-177)."""
    if order < 0:
        raise ValueError("order must be >= 0")
    prev, cur = 1, 2 * X
    if order == 0:
        return prev
    for _ in range(order - 1):
        prev, cur = cur, 2 * X * cur - prev
    return cur


def cheb1der(X, order):
    """dT_order/dX = order * U_{order-1} (pxmcmc/utils.py:180-197)."""
    if order < 0:
        raise ValueError("order must be > 0")
    return 0 if order == 0 else order * chebyshev2(X, order - 1)


# ------------------------------------------------------------------ MW sampling geometry
def mw_sample_positions(L):
    """(thetas[L], phis[2L-1]) of MW sampling (pyssht.sample_positions)."""
    n = 2 * L - 1
    return (2.0 * np.arange(L) + 1.0) * np.pi / n, 2.0 * np.pi * np.arange(n) / n


def mw_weights(m):
    """int_0^pi exp(i m theta) sin(theta) dtheta (pxmcmc/utils.py:249-259)."""
    if m == 1:
        return 1j * np.pi / 2
    if m == -1:
        return -1j * np.pi / 2
    if m % 2 == 0:
        return 2.0 / (1.0 - m * m)
    return 0


def weights_theta(L):
    """pxmcmc/utils.py:262-267."""
    n = 2 * L - 1
    ms = np.arange(-(L - 1), L)
    w = np.array([mw_weights(int(m)) for m in ms], dtype=complex) * np.exp(-1j * ms * np.pi / n)
    return (np.fft.fft(np.fft.ifftshift(w)) * 2 * np.pi / n ** 2).real


def mw_map_weights(L):
    """Exact MW quadrature weights as a flat (L(2L-1),) map (pxmcmc/utils.py:270-283)."""
    wr = weights_theta(L)
    q = wr[:L].copy()
    q[: L - 1] += wr[: L - 1: -1]
    return np.repeat(q, 2 * L - 1)


def s2_integrate(f, L):
    """Integral of an MW-sampled map over the sphere (pxmcmc/utils.py:286-299)."""
    return (mw_map_weights(L) * f).sum()


def pixel_area(r, theta1, theta2, phi1, phi2):
    return r ** 2 * (np.cos(theta1) - np.cos(theta2)) * (phi2 - phi1)


def polar_cap_area(r, theta):
    return 2 * np.pi * r ** 2 * (1 - np.cos(theta))


def calc_pixel_areas(L, r=1):
    """Areas of the MW pixels, shape (L, 2L-1) (pxmcmc/utils.py:227-246)."""
    thetas, phis = mw_sample_positions(L)
    nphi = phis.size
    dphi = np.diff(np.append(phis, 2 * np.pi))
    areas = np.empty((L, nphi))
    areas[0] = polar_cap_area(r, thetas[0]) / nphi
    band = r ** 2 * (np.cos(thetas[:-1]) - np.cos(thetas[1:]))
    areas[1:] = band[:, None] * dphi[None, :]
    return areas


def norm(x):
    return np.linalg.norm(x)


def snr(signal, noise):
    return 20 * np.log10(norm(signal) / norm(noise))


# ------------------------------------------------------------------ wavelet tiling
def wavelet_tiling(B, L, N, J_min, spin):
    """pys2let.wavelet_tiling for N=1, spin=0: (phi_l [L], psi_lm [L*L, nscales])."""
    if N != 1 or spin != 0:
        raise NotImplementedError("only axisymmetric (N=1), spin-0 wavelets are implemented")
    _, k0, k = _lib.wavelet_tiling_host(L, B, J_min)
    ls = np.arange(L)
    phi_l = (np.sqrt((2 * ls + 1) / (4 * np.pi)) * k0).astype(complex)
    psi_lm = np.zeros((L * L, k.shape[0]), dtype=complex)
    psi_lm[ls * ls + ls, :] = (np.sqrt((2 * ls + 1) / (8 * np.pi ** 2))[None, :] * k).T
    return phi_l, psi_lm


def j_max(L, B):
    return int(np.ceil(np.log(L) / np.log(B)))


def _multires_bandlimits(L, B, J_min, dirs=1, spin=0):
    """Bandlimit of the scaling function and of every wavelet scale, recovered
    from the support of the harmonic kernels (pxmcmc/utils.py:116-125)."""
    phi_l, psi_lm = wavelet_tiling(B, L, dirs, J_min, spin)
    ls = np.arange(L)
    rows = [phi_l] + [psi_lm[ls * ls + ls, j] for j in range(psi_lm.shape[1])]
    return np.array([np.nonzero(r)[0].max() + 1 for r in rows], dtype=int)


def mw_size(L):
    return L * (2 * L - 1)


def map2alm(image, lmax, **kwargs):
    raise NotImplementedError("HEALPix map2alm is setup-time only and not part of this build yet (SURVEY.md 8 a13)")


def alm2map(alm, nside, **kwargs):
    raise NotImplementedError("HEALPix alm2map is setup-time only and not part of this build yet (SURVEY.md 8 a13)")
