"""numpy restatement of the s2let (+so3, N=1) scale-discretised wavelet
transform on MW sampling.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the published algorithm of s2let (Leistedt et al. 2013; McEwen et al.
2015) as used through ``pys2let==2.2.6`` at the reference's call sites
``pxmcmc/transforms.py:75,95-98,164``, ``pxmcmc/prior.py:72,121,132``,
``pxmcmc/utils.py:117`` and ``pxmcmc/forward.py:1,109`` with the parameters the
reference fixes: ``N=1`` (axisymmetric), ``spin=0``, ``upsample=0``
(multiresolution).  The wheel is absent => *parity unpinned* against the real
binary; pinned by partition of unity, round trip and the adjoint dot-tests of
``tests/test_transforms.py:16-46``.

Maths:
  J = ceil(log L / log B);  L_j = min(ceil(B^(j+1)), L);  L_s = min(ceil(B^J_min), L)
  phi2_j(l) = 1 (l < B^(j-1)), 0 (l > B^j), else Q(l/B^j, 1)/Q(1/B, 1) with Q the
  300-step trapezoid of exp(-2/(1-t^2))/k, t = (k - 1/B) 2B/(B-1) - 1
  kappa0 = sqrt(phi2_{J_min}),  kappa_j = sqrt(phi2_{j+1} - phi2_j)
  analysis : S = A_inv(L_s)[kappa0 flm],  W_j = (2 pi)^-1/2 A_inv(L_j)[kappa_j flm]
  synthesis: flm = kappa0 A_fwd(L_s) S + sum_j (2 pi)^1/2 kappa_j A_fwd(L_j) W_j
  (the (2 pi)^-+1/2 pair is so3's N=1 normalisation), adjoints = exact conjugate
  transposes.
"""
from math import ceil, exp, isfinite, log, pi, sqrt

import numpy as np

from . import ssht_ref


def j_max(L, B):
    return int(ceil(log(L) / log(B)))


def mw_size(L):
    return L * (2 * L - 1)


def _f_s2dw(k, B):
    t = (k - (1.0 / B)) * (2.0 * B / (B - 1.0)) - 1.0
    den = 1.0 - t * t
    if den == 0.0:
        return float("nan")
    try:
        return exp(-2.0 / den) / k
    except OverflowError:
        return float("inf")


def _quadtrap(a, b, n, B):
    if a == b:
        return 0.0
    h = (b - a) / n
    tot = 0.0
    for i in range(n):
        f1 = _f_s2dw(a + i * h, B)
        f2 = _f_s2dw(a + (i + 1) * h, B)
        if isfinite(f1) and isfinite(f2):
            tot += ((f1 + f2) * h) / 2.0
    return tot


def tiling_axisym(B, L, J_min):
    """kappa[j, l] for j = 0..J (rows j < J_min zero) and kappa0[l]."""
    J = j_max(L, B)
    n = 300
    norm = _quadtrap(1.0 / B, 1.0, n, B)
    phi2 = np.zeros((J + 2, L))
    for j in range(J + 2):
        for l in range(L):
            if l < B ** (j - 1):
                phi2[j, l] = 1.0
            elif l > B ** j:
                phi2[j, l] = 0.0
            else:
                phi2[j, l] = _quadtrap(l / B ** j, 1.0, n, B) / norm
    kappa0 = np.sqrt(phi2[J_min])
    kappa = np.zeros((J + 1, L))
    for j in range(J_min, J + 1):
        diff = phi2[j + 1] - phi2[j]
        kappa[j] = np.sqrt(np.clip(diff, 0.0, None))
    return kappa, kappa0


def bandlimits(B, L, J_min):
    """[L_s, L_{J_min}, ..., L_J] (scaling function first)."""
    J = j_max(L, B)
    out = [min(int(ceil(B ** J_min)), L)]
    for j in range(J_min, J + 1):
        out.append(min(int(ceil(B ** (j + 1))), L))
    return out


def wavelet_tiling(B, L, N, J_min, spin):
    """pys2let.wavelet_tiling restated for N=1, spin=0: returns (phi_l [L],
    psi_lm [L^2, nscales]) with phi_l = sqrt((2l+1)/4pi) kappa0 and
    psi^j_{l0} = sqrt((2l+1)/8pi^2) kappa_j (m != 0 entries zero)."""
    assert N == 1 and spin == 0
    kappa, kappa0 = tiling_axisym(B, L, J_min)
    J = j_max(L, B)
    ls = np.arange(L)
    phi_l = (np.sqrt((2 * ls + 1) / (4 * pi)) * kappa0).astype(complex)
    psi_lm = np.zeros((L * L, J - J_min + 1), dtype=complex)
    for j in range(J_min, J + 1):
        psi_lm[ls * ls + ls, j - J_min] = np.sqrt((2 * ls + 1) / (8 * pi * pi)) * kappa[j]
    return phi_l, psi_lm


def harmonic_kernels(B, L, J_min):
    """Per-scale harmonic multipliers, scaling first: list of (L_j, g_j[l]) with
    g_s = kappa0 and g_j = kappa_j  (the sqrt(2 pi) factors are applied by the
    callers below so that analysis and synthesis read like A.5)."""
    kappa, kappa0 = tiling_axisym(B, L, J_min)
    J = j_max(L, B)
    bl = bandlimits(B, L, J_min)
    out = [(bl[0], kappa0[: bl[0]].copy())]
    for j in range(J_min, J + 1):
        Lj = bl[1 + j - J_min]
        out.append((Lj, kappa[j, :Lj].copy()))
    return out


def _per_l(g, Lj):
    """expand g[l] to the l^2+l+m layout at bandlimit Lj."""
    return np.repeat(g[:Lj], 2 * np.arange(Lj) + 1)


_SQ2PI = sqrt(2.0 * pi)


def analysis_px2wav(f, B, L, J_min, N=1, spin=0, upsample=0):
    """f (L(2L-1),) -> (f_wav, f_scal), multiresolution."""
    assert N == 1 and spin == 0 and upsample == 0
    flm = ssht_ref.forward(np.asarray(f, dtype=complex).reshape(L, 2 * L - 1), L, 0)
    ks = harmonic_kernels(B, L, J_min)
    Ls, g = ks[0]
    f_scal = ssht_ref.inverse(flm[: Ls * Ls] * _per_l(g, Ls), Ls, 0).ravel()
    wavs = []
    for Lj, g in ks[1:]:
        wavs.append(ssht_ref.inverse(flm[: Lj * Lj] * _per_l(g, Lj) / _SQ2PI, Lj, 0).ravel())
    return np.concatenate(wavs), f_scal


def synthesis_wav2px(f_wav, f_scal, B, L, J_min, N=1, spin=0, upsample=0):
    assert N == 1 and spin == 0 and upsample == 0
    ks = harmonic_kernels(B, L, J_min)
    flm = np.zeros(L * L, dtype=complex)
    Ls, g = ks[0]
    flm[: Ls * Ls] += _per_l(g, Ls) * ssht_ref.forward(np.asarray(f_scal, dtype=complex).reshape(Ls, 2 * Ls - 1), Ls, 0)
    off = 0
    f_wav = np.asarray(f_wav, dtype=complex)
    for Lj, g in ks[1:]:
        nj = mw_size(Lj)
        flm[: Lj * Lj] += _SQ2PI * _per_l(g, Lj) * ssht_ref.forward(f_wav[off: off + nj].reshape(Lj, 2 * Lj - 1), Lj, 0)
        off += nj
    assert off == f_wav.size
    return ssht_ref.inverse(flm, L, 0).ravel()


def synthesis_adjoint_px2wav(f, B, L, J_min, N=1, spin=0, upsample=0):
    """Exact conjugate transpose of synthesis_wav2px."""
    assert N == 1 and spin == 0 and upsample == 0
    flm = ssht_ref.inverse_adjoint(np.asarray(f, dtype=complex).reshape(L, 2 * L - 1), L, 0)
    ks = harmonic_kernels(B, L, J_min)
    Ls, g = ks[0]
    f_scal = ssht_ref.forward_adjoint(flm[: Ls * Ls] * _per_l(g, Ls), Ls, 0).ravel()
    wavs = []
    for Lj, g in ks[1:]:
        wavs.append(ssht_ref.forward_adjoint(_SQ2PI * flm[: Lj * Lj] * _per_l(g, Lj), Lj, 0).ravel())
    return np.concatenate(wavs), f_scal


def analysis_adjoint_wav2px(f_wav, f_scal, B, L, J_min, N=1, spin=0, upsample=0):
    """Exact conjugate transpose of analysis_px2wav."""
    assert N == 1 and spin == 0 and upsample == 0
    ks = harmonic_kernels(B, L, J_min)
    flm = np.zeros(L * L, dtype=complex)
    Ls, g = ks[0]
    flm[: Ls * Ls] += _per_l(g, Ls) * ssht_ref.inverse_adjoint(np.asarray(f_scal, dtype=complex).reshape(Ls, 2 * Ls - 1), Ls, 0)
    off = 0
    f_wav = np.asarray(f_wav, dtype=complex)
    for Lj, g in ks[1:]:
        nj = mw_size(Lj)
        flm[: Lj * Lj] += _per_l(g, Lj) / _SQ2PI * ssht_ref.inverse_adjoint(f_wav[off: off + nj].reshape(Lj, 2 * Lj - 1), Lj, 0)
        off += nj
    assert off == f_wav.size
    return ssht_ref.forward_adjoint(flm, L, 0).ravel()


def lm_hp2lm(alm_hp, L):
    """healpy m-major (m >= 0) alm -> ssht l^2+l+m layout."""
    flm = np.zeros(L * L, dtype=complex)
    lmax = L - 1
    for m in range(L):
        for l in range(m, L):
            a = alm_hp[m * (2 * lmax + 1 - m) // 2 + l]
            flm[l * l + l + m] = a
            if m > 0:
                flm[l * l + l - m] = (-1.0) ** m * np.conj(a)
    return flm


def alm2map_mw(flm, L, spin):
    return ssht_ref.inverse(flm, L, spin).ravel()
