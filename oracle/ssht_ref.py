"""numpy restatement of the ssht spin spherical harmonic transforms on
McEwen-Wiaux (MW) sampling.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the published algorithm of ssht (McEwen & Wiaux 2011, "A novel sampling
theorem on the sphere"), the C library behind ``pyssht==1.5.2`` which is what the
reference calls at ``pxmcmc/measurements.py:223,225,237,239`` and, through
s2let, in every wavelet transform (``pxmcmc/transforms.py:95-98``).  The wheel
is absent from the image => against the real binary this file is *parity
unpinned*; it is pinned by closed-form harmonics, exactness, and the
adjoint/round-trip properties the reference's tests assert
(``tests/test_transforms.py:16-46``, ``tests/test_utils.py:85-100``,
``tests/test_measurements.py:71-130``).

Algorithm (follows ssht_core_mw_{inverse,forward}_sov[_conv]_sym and
ssht_adjoint_mw_*): Wigner functions are expanded in the Fourier basis through
their values at pi/2,

    d^l_{mn}(theta) = i^(n-m) sum_{m'} D^l_{m'm} D^l_{m'n} exp(i m' theta),
    D^l = d^l(pi/2),

so every transform is a contraction over l with D^l planes plus DFTs in theta
and phi.  This is deliberately a different route from the GPU path (per-ring
Wigner recurrences + per-m dense contractions).

Conventions:
  theta_t = (2t+1) pi/(2L-1), t=0..L-1 ; phi_p = 2 pi p/(2L-1), p=0..2L-2
  f is (L, 2L-1) row-major [t, p];  flm has L^2 entries, index l^2+l+m
  sYlm(theta,phi) = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta) exp(i m phi)
"""
from math import lgamma, log, pi

import numpy as np


# --------------------------------------------------------------------------
# sampling helpers (pyssht.sample_* / elm2ind equivalents)
# --------------------------------------------------------------------------
def sample_length(L):
    return L * (2 * L - 1)


def sample_shape(L):
    return (L, 2 * L - 1)


def sample_positions(L):
    n = 2 * L - 1
    thetas = (2.0 * np.arange(L) + 1.0) * pi / n
    phis = 2.0 * pi * np.arange(n) / n
    return thetas, phis


def elm2ind(el, m):
    return el * el + el + m


def ind2elm(ind):
    el = int(np.floor(np.sqrt(ind)))
    return el, ind - el * el - el


def theta_to_index(theta, L):
    # pyssht.theta_to_index: nearest-below MW ring
    return int(np.floor((theta * (2 * L - 1) / pi - 1.0) / 2.0 + 0.5))


def phi_to_index(phi, L):
    return int(np.floor(phi * (2 * L - 1) / (2 * pi) + 0.5)) % (2 * L - 1)


def mw_quad_weight(m):
    """w(m) = int_0^pi exp(i m theta) sin(theta) dtheta."""
    if m == 1:
        return 0.5j * pi
    if m == -1:
        return -0.5j * pi
    if m % 2 == 0:
        return 2.0 / (1.0 - m * m)
    return 0.0


# --------------------------------------------------------------------------
# Wigner d at pi/2, one l-plane at a time
# --------------------------------------------------------------------------
def wigner_pi2_planes(L):
    """Yield (l, D) for l = 0..L-1 with D[mp, m+L-1] = d^l_{mp,m}(pi/2) for
    0 <= mp <= l, |m| <= l (zero outside).  Three-term recurrence in l at
    fixed (mp, m), seeded on the boundary max(mp,|m|) = l by the closed form
    d^l_{l,n}(pi/2) = (-1)^(l-n) sqrt((2l)!/((l+n)!(l-n)!)) / 2^l.
    """
    n = 2 * L - 1
    off = L - 1
    mp = np.arange(L, dtype=np.float64)[:, None]
    mm = np.arange(-(L - 1), L, dtype=np.float64)[None, :]
    Dm1 = np.zeros((L, n))
    D0 = np.zeros((L, n))
    D0[0, off] = 1.0
    yield 0, D0.copy()
    for l in range(0, L - 1):
        l1 = l + 1
        Dn = np.zeros((L, n))
        if l >= 1:
            a = np.sqrt(np.clip((l * l - mp * mp) * (l * l - mm * mm), 0.0, None))
            b = np.sqrt(np.clip((l1 * l1 - mp * mp) * (l1 * l1 - mm * mm), 0.0, None))
            sl = (slice(0, l + 1), slice(off - l, off + l + 1))
            num = (2 * l + 1) * (-(mp * mm))[sl] * D0[sl] - l1 * a[sl] * Dm1[sl]
            Dn[sl] = num / (l * b[sl])
        # (l = 0: d^1_{00}(pi/2) = cos(pi/2) = 0, already zero)
        # boundary seeds: row mp = l1
        ns = np.arange(-l1, l1 + 1)
        lg = np.array(
            [0.5 * (lgamma(2 * l1 + 1) - lgamma(l1 + k + 1) - lgamma(l1 - k + 1)) for k in ns]
        ) - l1 * log(2.0)
        seed = np.where((l1 - ns) % 2 == 0, 1.0, -1.0) * np.exp(lg)  # d^{l1}_{l1,n}
        Dn[l1, off - l1: off + l1 + 1] = seed
        # columns m = +l1 and m = -l1 for mp < l1
        mps = np.arange(0, l1)
        # d_{mp,l1} = (-1)^(mp-l1) d_{l1,mp}
        Dn[mps, off + l1] = np.where((l1 - mps) % 2 == 0, 1.0, -1.0) * seed[mps + l1]
        # d_{mp,-l1} = d_{l1,-mp}
        Dn[mps, off - l1] = seed[-mps + l1]
        Dm1, D0 = D0, Dn
        yield l1, D0.copy()


def wigner_d_explicit(l, m, n, beta):
    """Closed-form d^l_{mn}(beta); small l only (factorials)."""
    from math import factorial, cos, sin, sqrt

    pref = sqrt(factorial(l + n) * factorial(l - n) * factorial(l + m) * factorial(l - m))
    tot = 0.0
    for k in range(max(0, n - m), min(l + n, l - m) + 1):
        den = factorial(l + n - k) * factorial(k) * factorial(l - k - m) * factorial(k - n + m)
        tot += (
            (-1.0) ** (k - n + m)
            / den
            * cos(beta / 2) ** (2 * l - 2 * k + n - m)
            * sin(beta / 2) ** (2 * k - n + m)
        )
    return pref * tot


def sylm_explicit(s, l, m, theta, phi):
    """Spin-weighted spherical harmonic from the closed-form Wigner d."""
    return (
        (-1.0) ** s
        * np.sqrt((2 * l + 1) / (4 * pi))
        * wigner_d_explicit(l, m, -s, theta)
        * np.exp(1j * m * phi)
    )


# --------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------
def _phases(L):
    n = 2 * L - 1
    ms = np.arange(-(L - 1), L)
    thetas, phis = sample_positions(L)
    theta_ext = (2.0 * np.arange(n) + 1.0) * pi / n
    Ephi = np.exp(1j * np.outer(ms, phis))  # [m, p]
    Eth = np.exp(1j * np.outer(ms, thetas))  # [m', t<L]
    Eth_ext = np.exp(1j * np.outer(ms, theta_ext))  # [m', t<2L-1]
    return ms, Ephi, Eth, Eth_ext


def _im_pow(k):
    """i**k for integer arrays (exact)."""
    return np.array([1, 1j, -1, -1j])[np.mod(k, 4)]


def _contract_to_lm(G, L, spin):
    """flm[l,m] = (-1)^s c_l i^-(m+s) sum_{m'=-l..l} D^l_{m'm} D^l_{m',-s} G[m',m].

    G is (2L-1, 2L-1) indexed [m'+L-1, m+L-1]."""
    off = L - 1
    ms = np.arange(-(L - 1), L)
    par = np.where((ms + spin) % 2 == 0, 1.0, -1.0)[None, :]  # (-1)^(m+s)
    # fold negative m' onto m' >= 0:  D_{-m',m} D_{-m',-s} = (-1)^(m+s) D_{m'm} D_{m',-s}
    Gh = G[off:, :].copy()
    Gh[1:, :] += par * G[off - 1:: -1, :][: L - 1]
    flm = np.zeros(L * L, dtype=complex)
    ssign = -1.0 if spin % 2 else 1.0
    for l, D in wigner_pi2_planes(L):
        if l < abs(spin):
            continue
        c = ssign * np.sqrt((2 * l + 1) / (4 * pi))
        cols = slice(off - l, off + l + 1)
        dsp = D[: l + 1, off - spin]
        val = np.einsum("am,a,am->m", D[: l + 1, cols], dsp, Gh[: l + 1, cols])
        mloc = np.arange(-l, l + 1)
        flm[l * l + l + mloc] = c * _im_pow(-(mloc + spin)) * val
    return flm


def _expand_from_lm(flm, L, spin, conj_kernel=False):
    """Fmm[m',m] = sum_l K[l,m',m] flm[l,m] with
    K = (-1)^s c_l i^-(m+s) D^l_{m'm} D^l_{m',-s}  (or its complex conjugate)."""
    n = 2 * L - 1
    off = L - 1
    ms = np.arange(-(L - 1), L)
    par = np.where((ms + spin) % 2 == 0, 1.0, -1.0)[None, :]
    Fh = np.zeros((L, n), dtype=complex)
    ssign = -1.0 if spin % 2 else 1.0
    for l, D in wigner_pi2_planes(L):
        if l < abs(spin):
            continue
        c = ssign * np.sqrt((2 * l + 1) / (4 * pi))
        cols = slice(off - l, off + l + 1)
        mloc = np.arange(-l, l + 1)
        ph = _im_pow(-(mloc + spin))
        if conj_kernel:
            ph = np.conj(ph)
        coef = c * ph * flm[l * l + l + mloc]
        dsp = D[: l + 1, off - spin]
        Fh[: l + 1, cols] += D[: l + 1, cols] * dsp[:, None] * coef[None, :]
    Fmm = np.zeros((n, n), dtype=complex)
    Fmm[off:, :] = Fh
    Fmm[:off, :] = (par * Fh[1:, :])[::-1, :]
    return Fmm


# --------------------------------------------------------------------------
# the four primitives
# --------------------------------------------------------------------------
def inverse(flm, L, spin=0):
    """flm -> f (L, 2L-1).  ssht_core_mw_inverse_sov_sym."""
    flm = np.asarray(flm, dtype=complex).ravel()
    assert flm.size == L * L
    _, Ephi, Eth, _ = _phases(L)
    Fmm = _expand_from_lm(flm, L, spin)
    return Eth.T @ Fmm @ Ephi


def inverse_adjoint(f, L, spin=0):
    """f -> flm, exact Euclidean adjoint of `inverse`.  ssht_adjoint_mw_inverse_sov_sym."""
    f = np.asarray(f, dtype=complex).reshape(L, 2 * L - 1)
    _, Ephi, Eth, _ = _phases(L)
    Fm = f @ np.conj(Ephi).T  # [t, m]
    H = Eth @ Fm  # [m', m] = sum_t e^{i m' theta_t} Fm[t,m]
    return _contract_to_lm(H, L, spin)


def _wconv(L):
    ms = np.arange(-(L - 1), L)
    return np.array([[mw_quad_weight(int(a + b)) for b in ms] for a in ms], dtype=complex)


def forward(f, L, spin=0):
    """f -> flm with the exact MW quadrature.  ssht_core_mw_forward_sov_conv_sym."""
    f = np.asarray(f, dtype=complex).reshape(L, 2 * L - 1)
    n = 2 * L - 1
    ms, Ephi, _, Eth_ext = _phases(L)
    par = np.where((ms + spin) % 2 == 0, 1.0, -1.0)[None, :]
    Fm = f @ np.conj(Ephi).T / n  # [t, m]
    Fext = np.zeros((n, n), dtype=complex)
    Fext[:L] = Fm
    Fext[L:] = (par * Fm[: L - 1])[::-1]  # t -> 2L-2-t
    Fmm = np.conj(Eth_ext) @ Fext / n  # [m'', m]
    G = 2 * pi * (_wconv(L) @ Fmm)  # [m', m]
    return _contract_to_lm(G, L, spin)


def forward_adjoint(flm, L, spin=0):
    """flm -> f, exact Euclidean adjoint of `forward`.  ssht_adjoint_mw_forward_sov_sym."""
    flm = np.asarray(flm, dtype=complex).ravel()
    n = 2 * L - 1
    ms, Ephi, _, Eth_ext = _phases(L)
    par = np.where((ms + spin) % 2 == 0, 1.0, -1.0)[None, :]
    Gd = _expand_from_lm(flm, L, spin, conj_kernel=True)  # [m', m]
    Fmm = 2 * pi * (np.conj(_wconv(L)).T @ Gd)
    Fext = Eth_ext.T @ Fmm / n  # [t<2L-1, m]
    Fm = Fext[:L].copy()
    Fm[: L - 1] += par * Fext[L:][::-1]
    return Fm @ Ephi / n


# --------------------------------------------------------------------------
# slow independent references used only to pin the above at small L
# --------------------------------------------------------------------------
def inverse_direct(flm, L, spin=0):
    """Direct summation over closed-form sYlm (O(L^4), L <= ~12)."""
    thetas, phis = sample_positions(L)
    f = np.zeros((L, 2 * L - 1), dtype=complex)
    for l in range(abs(spin), L):
        for m in range(-l, l + 1):
            c = flm[l * l + l + m]
            if c == 0:
                continue
            for t, th in enumerate(thetas):
                f[t] += c * sylm_explicit(spin, l, m, th, phis)
    return f


def forward_quadrature_matrix(L, spin, m):
    """Dense (L-|m|... ) operator of `forward` for one m, built the SURVEY.md A.2
    way: trigonometric interpolation of the theta-extended ring samples onto a
    bandlimit-L' >= 2L-1 MW grid, then the exact MW quadrature there.  Returns
    W[l, t] (l = 0..L-1, rows l < max(|m|,|s|) zero) such that
    flm[l,m] = sum_t W[l,t] F_m(theta_t),  F_m = (1/(2L-1)) sum_p f e^{-i m phi_p}."""
    n = 2 * L - 1
    Lp = 2 * L
    npp = 2 * Lp - 1
    mpp = np.arange(-(L - 1), L)
    theta_ext = (2.0 * np.arange(n) + 1.0) * pi / n
    th_f = (2.0 * np.arange(Lp) + 1.0) * pi / npp
    par = 1.0 if (m + spin) % 2 == 0 else -1.0
    R = np.zeros((n, L))
    R[:L] = np.eye(L)
    for t in range(L - 1):
        R[n - 1 - t, t] = par
    dft = np.exp(-1j * np.outer(mpp, theta_ext)) / n  # [m'', t]
    th_f_ext = (2.0 * np.arange(npp) + 1.0) * pi / npp
    interp = (np.exp(1j * np.outer(th_f_ext, mpp)) @ dft @ R).real  # [k < 2L'-1, t]
    # quadrature weights over the full theta-extended fine grid (theta part only).
    # NOTE: the interpolant of an arbitrary south-pole sample does not have the
    # (-1)^(m+s) reflection parity, so the folded weights of utils.mw_map_weights
    # are not enough here; fold the *product* weights x interpolant instead.
    wr = np.array([mw_quad_weight(int(k)) * np.exp(-1j * k * pi / npp) for k in range(-(Lp - 1), Lp)])
    wr = (np.fft.fft(np.fft.ifftshift(wr)) / npp).real
    Q = wr[:Lp, None] * interp[:Lp]
    Q[: Lp - 1] += par * (wr[npp - 1: Lp - 1: -1, None] * interp[npp - 1: Lp - 1: -1])
    W = np.zeros((L, L))
    for l in range(max(abs(m), abs(spin)), L):
        lam = np.array(
            [(-1.0) ** spin * np.sqrt((2 * l + 1) / (4 * pi)) * wigner_d_explicit(l, m, -spin, th) for th in th_f]
        )
        W[l] = 2 * pi * lam @ Q
    return W
