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


# ----------------------------------------------------------------------------
# HEALPix <-> harmonic space (data preparation of the reference's drivers)
# ----------------------------------------------------------------------------
class _LazyDevice:
    """pxmcmc_b200.device on first use (keeps `import pxmcmc_b200.utils` free of torch)"""

    def __getattr__(self, name):
        from . import device

        return getattr(device, name)


D = _LazyDevice()


def alm_hp_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def alm_hp_index(el, m, lmax):
    """healpy.Alm.getidx: m-major storage of the m >= 0 coefficients"""
    return m * (2 * lmax + 1 - m) // 2 + el


def lm_hp2lm(flm_hp, L):
    """pys2let.lm_hp2lm (experiments/earthtopography/main.py:82): healpy's m >= 0 alm of a REAL map ->
    ssht ordering l*l+l+m with f_{l,-m} = (-1)^m conj(f_{l,m})"""
    flm_hp = np.asarray(flm_hp, dtype=complex)
    lmax = L - 1
    if flm_hp.size != alm_hp_size(lmax):
        raise ValueError("alm has the wrong size for this bandlimit")
    flm = np.zeros(L * L, dtype=complex)
    for m in range(L):
        els = np.arange(m, L)
        a = flm_hp[alm_hp_index(els, m, lmax)]
        flm[els * els + els + m] = a
        if m > 0:
            flm[els * els + els - m] = (-1.0) ** m * np.conj(a)
    return flm


def lm2lm_hp(flm, L):
    """pys2let.lm2lm_hp: the m >= 0 half in healpy's storage"""
    flm = np.asarray(flm, dtype=complex)
    lmax = L - 1
    out = np.zeros(alm_hp_size(lmax), dtype=complex)
    for m in range(L):
        els = np.arange(m, L)
        out[alm_hp_index(els, m, lmax)] = flm[els * els + els + m]
    return out


def alm2map_mw(flm, L, spin=0):
    """pys2let.alm2map_mw = ssht inverse transform on MW sampling; returns the flat complex map
    (experiments/earthtopography/main.py:82, tests/conftest.py:49 of the reference)"""
    from . import sht

    return np.asarray(sht.inverse(flm, L, Spin=spin)).ravel()


class _HealpixPlan:
    _cache = {}

    @classmethod
    def get(cls, nside, L):
        import torch

        key = (int(nside), int(L), torch.cuda.current_device() if torch.cuda.is_available() else -1)
        if key not in cls._cache:
            cls._cache[key] = cls(nside, L)
        return cls._cache[key]

    def __init__(self, nside, L):
        import ctypes as C

        from ._lib import check, lib

        D.dev()
        self.nside, self.L, self.npix = int(nside), int(L), 12 * int(nside) ** 2
        h = C.c_void_p()
        check(lib.pxm_hpx_plan_create(self.nside, self.L, C.byref(h)))
        self.h = h

    def synth(self, flm_d):
        """complex device flm [L*L] -> complex device map [npix]"""
        import torch

        from ._lib import check, lib, ptr, stream_ptr

        out = torch.empty(self.npix, dtype=D.CDT, device=flm_d.device)
        check(lib.pxm_hpx_alm2map(self.h, ptr(flm_d), ptr(out), stream_ptr()))
        return out

    def adjoint(self, map_d):
        import torch

        from ._lib import check, lib, ptr, stream_ptr

        out = torch.empty(self.L * self.L, dtype=D.CDT, device=map_d.device)
        check(lib.pxm_hpx_map2alm_adjoint(self.h, ptr(map_d), ptr(out), stream_ptr()))
        return out


def _nside_of(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix or nside & (nside - 1):
        raise ValueError("not a HEALPix map: size must be 12 nside^2 with nside a power of two")
    return nside


def map2alm(image, lmax, iter=3, **kwargs):
    """healpy.map2alm as the reference wraps it (pxmcmc/utils.py:106-108): RING-ordered real map ->
    a_lm, m >= 0, healpy storage; uniform weights 4 pi / npix and `iter` Jacobi refinements
    (healpy's default iter=3).  Every transform runs on the device (pxm_hpx_*)."""
    unsupported = {k: v for k, v in kwargs.items() if k not in ("mmax", "pol", "use_weights", "use_pixel_weights", "verbose", "datapath", "gal_cut")}
    if unsupported or kwargs.get("use_weights") or kwargs.get("use_pixel_weights") or kwargs.get("gal_cut"):
        raise NotImplementedError(f"map2alm options not supported: {sorted(kwargs)}")
    if kwargs.get("mmax") not in (None, lmax):
        raise NotImplementedError("mmax != lmax is not supported")
    image = np.asarray(image, dtype=float)
    if image.ndim != 1:
        raise NotImplementedError("polarised (T, Q, U) maps are not supported")
    L = int(lmax) + 1
    plan = _HealpixPlan.get(_nside_of(image.size), L)
    w = 4.0 * np.pi / plan.npix
    f = D.to_dev_c(image)
    flm = D.lincomb_dev([(w, plan.adjoint(f))])
    for _ in range(int(iter)):
        resid = D.lincomb_dev([(1.0, f), (-1.0, plan.synth(flm))])
        flm = D.lincomb_dev([(1.0, flm), (w, plan.adjoint(resid))])
    return lm2lm_hp(D.to_host(flm), L)


def alm2map(alm, nside, **kwargs):
    """healpy.alm2map as the reference wraps it (pxmcmc/utils.py:111-113): a_lm (m >= 0, healpy
    storage, lmax inferred from the size) -> RING-ordered real map"""
    unsupported = {k for k in kwargs if k not in ("lmax", "mmax", "pol", "verbose", "inplace")}
    if unsupported:
        raise NotImplementedError(f"alm2map options not supported: {sorted(unsupported)}")
    alm = np.asarray(alm, dtype=complex)
    if alm.ndim != 1:
        raise NotImplementedError("polarised alm are not supported")
    lmax = kwargs.get("lmax")
    if lmax is None:
        lmax = int(round((-3 + np.sqrt(1 + 8 * alm.size)) / 2))
    if alm_hp_size(lmax) != alm.size:
        raise ValueError("alm size does not correspond to an lmax (mmax = lmax)")
    L = lmax + 1
    plan = _HealpixPlan.get(int(nside), L)
    return D.to_host(plan.synth(D.to_dev_c(lm_hp2lm(alm, L)))).real.copy()


def build_mask(L, size=20):
    """galactic-plane + ecliptic mask on the MW grid (pxmcmc/utils.py:320-349); see ``ingest.build_mask``"""
    from .ingest import build_mask as _bm

    return _bm(L, size)
