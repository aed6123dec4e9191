"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the HEALPix leg of the data preparation.

Restates what the reference obtains from healpy 1.16.5 (pinned in
/root/reference/poetry.lock:606-608; wheel absent, no network) through
/root/reference/pxmcmc/utils.py:106-113 (`map2alm`, `alm2map`) and from pys2let's
`lm_hp2lm` (/root/reference/experiments/earthtopography/main.py:80-82):

* RING-scheme pixel centres (HEALPix primer, Gorski et al. 2005, eqs. 2-9);
* `alm2map`: f(p) = sum_l [a_l0 Y_l0(p) + 2 Re sum_{m>0} a_lm Y_lm(p)] for a real map;
* `map2alm`: a_lm = (4 pi / npix) sum_p f(p) conj(Y_lm(p)), then `iter` (default 3) Jacobi
  refinements a += map2alm_0(f - alm2map(a));
* healpy alm storage: index m (2 lmax + 1 - m) / 2 + l, m >= 0.

Route: the dense matrix Y[p, (l, m)] from scipy.special.sph_harm_y -- O(npix L^2), usable for
nside <= 16 -- deliberately unrelated to the GPU's ring/Legendre-recurrence route.
PARITY UNPINNED against the healpy binary (the reference's tests hold no golden HEALPix vector);
pinned instead by closed-form pixel centres, the monopole, and the properties tested in
tests/test_oracle_golden.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this.
"""
import numpy as np
from scipy.special import sph_harm_y


def nside2npix(nside):
    return 12 * nside * nside


def ring_table(nside):
    """per ring (north to south): n pixels, index of first pixel, phi shift flag, z = cos(theta)"""
    out = []
    npix, ncap = nside2npix(nside), 2 * nside * (nside - 1)
    for i in range(1, 4 * nside):
        if i < nside:
            out.append((4 * i, 2 * i * (i - 1), 1, 1.0 - i * i / (3.0 * nside * nside)))
        elif i <= 3 * nside:
            out.append((4 * nside, ncap + (i - nside) * 4 * nside, (i - nside + 1) % 2, 4.0 / 3.0 - 2.0 * i / (3.0 * nside)))
        else:
            ip = 4 * nside - i
            out.append((4 * ip, npix - 2 * ip * (ip + 1), 1, -(1.0 - ip * ip / (3.0 * nside * nside))))
    return out


def pix2ang(nside):
    """(theta, phi) of every pixel centre, RING order"""
    theta = np.zeros(nside2npix(nside))
    phi = np.zeros(nside2npix(nside))
    for n, start, shift, z in ring_table(nside):
        theta[start:start + n] = np.arccos(z)
        phi[start:start + n] = (np.arange(n) + 0.5 * shift) * 2.0 * np.pi / n
    return theta, phi


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def alm_index(el, m, lmax):
    return m * (2 * lmax + 1 - m) // 2 + el


def lm_hp2lm(alm, L):
    lmax = L - 1
    flm = np.zeros(L * L, dtype=complex)
    for el in range(L):
        for m in range(el + 1):
            a = alm[alm_index(el, m, lmax)]
            flm[el * el + el + m] = a
            if m:
                flm[el * el + el - m] = (-1) ** m * np.conj(a)
    return flm


def lm2lm_hp(flm, L):
    lmax = L - 1
    alm = np.zeros(alm_size(lmax), dtype=complex)
    for el in range(L):
        for m in range(el + 1):
            alm[alm_index(el, m, lmax)] = flm[el * el + el + m]
    return alm


def ylm_matrix(nside, L):
    """Y[p, l*l+l+m] for l < L, |m| <= l"""
    theta, phi = pix2ang(nside)
    Y = np.zeros((theta.size, L * L), dtype=complex)
    for el in range(L):
        for m in range(-el, el + 1):
            Y[:, el * el + el + m] = sph_harm_y(el, m, theta, phi)
    return Y


def synthesis_complex(flm, nside, L, Y=None):
    Y = ylm_matrix(nside, L) if Y is None else Y
    return Y @ np.asarray(flm, dtype=complex)


def adjoint_complex(f, nside, L, Y=None):
    Y = ylm_matrix(nside, L) if Y is None else Y
    return np.conj(Y).T @ np.asarray(f, dtype=complex)


def alm2map(alm, nside, Y=None):
    lmax = int(round((-3 + np.sqrt(1 + 8 * len(alm))) / 2))
    L = lmax + 1
    return synthesis_complex(lm_hp2lm(np.asarray(alm, dtype=complex), L), nside, L, Y).real


def map2alm(image, lmax, iter=3, Y=None):
    image = np.asarray(image, dtype=float)
    nside = int(round(np.sqrt(image.size / 12)))
    L = lmax + 1
    Y = ylm_matrix(nside, L) if Y is None else Y
    w = 4.0 * np.pi / image.size
    flm = w * adjoint_complex(image, nside, L, Y)
    for _ in range(iter):
        flm = flm + w * adjoint_complex(image - synthesis_complex(flm, nside, L, Y).real, nside, L, Y)
    return lm2lm_hp(flm, L)
