"""Uncertainty quantification mirroring ``pxmcmc/uncertainty.py`` of the reference
(credible-interval ranges of the saved chain, credible-region threshold).

The per-parameter quantiles run on the GPU (``pxm_quantile_columns``: every column of the stored
chain is sorted in shared memory); numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
``chain_to_pixels`` is the batched replacement of the per-sample ``transform.inverse`` loop of the
reference's plot scripts (experiments/earthtopography/plot.py:103-112).
"""
import numpy as np
import torch

from . import device as D
from ._lib import check, lib, ptr, stream_ptr
from .utils import _multires_bandlimits


def _virtual_index(n, q):
    """numpy's default "linear" method: lower order statistic and interpolation weight of the virtual
    index (n - 1) q, evaluated with numpy's own floating-point expressions
    (numpy/lib/_function_base_impl.py: _QuantileMethods['linear'], _get_indexes, _get_gamma)"""
    q = np.float64(q)
    if not 0.0 <= q <= 1.0:
        raise ValueError("Quantiles must be in the range [0, 1]")
    v = (n - 1) * q
    lo = np.floor(v)
    gamma = v - lo
    if v >= n - 1:  # at or above the last order statistic: numpy takes the maximum
        return n - 1, 0.0
    return int(lo), float(gamma)


def _quantile_pair_dev(chain_d, qa, qb):
    """chain_d: CUDA tensor [nsamples, ncols] float64 with unit column stride -> (qa-, qb-quantile)[ncols]"""
    n, ncols = chain_d.shape
    if chain_d.stride(1) != 1:
        chain_d = chain_d.contiguous()
    lo_a, g_a = _virtual_index(n, qa)
    lo_b, g_b = _virtual_index(n, qb)
    out_a = torch.empty(ncols, dtype=torch.float64, device=chain_d.device)
    out_b = torch.empty_like(out_a)
    check(lib.pxm_quantile_columns(ptr(chain_d), n, ncols, chain_d.stride(0), lo_a, g_a, lo_b, g_b, ptr(out_a), ptr(out_b),
                                   stream_ptr()))
    return out_a, out_b


def _host_range(chain, qa, qb):
    """what the reference itself evaluates (pxmcmc/uncertainty.py:14-16)"""
    return np.quantile(chain, qb, axis=0) - np.quantile(chain, qa, axis=0)


def credible_interval_range(chain, alpha=0.05, max_bytes=2 << 30):
    """Range of the (1 - alpha) credible interval of every parameter: the difference of the
    1 - alpha/2 and alpha/2 quantiles over the samples (pxmcmc/uncertainty.py:7-16).

    :param chain: nsamples x nparams real array (numpy, or a float64 CUDA tensor)
    :return: nparams ranges (same kind of array as ``chain``)"""
    qa, qb = alpha / 2, 1 - alpha / 2
    if D.is_dev(chain):
        if chain.is_complex():
            raise TypeError("a must be an array of real numbers")
        if chain.shape[0] > lib.pxm_quantile_columns_max_samples():
            # longer than the shared-memory sort holds: numpy's own quantile on the host (rare: the kernel takes 16 384
            # samples per parameter; the reference's default of 10^6 saved samples would not fit the GPU anyway)
            return torch.from_numpy(_host_range(D.to_host(chain), qa, qb)).to(chain.device)
        lo, hi = _quantile_pair_dev(chain.to(torch.float64), qa, qb)
        return hi - lo
    chain = np.asarray(chain)
    if np.iscomplexobj(chain):
        raise TypeError("a must be an array of real numbers")
    if chain.ndim != 2:
        raise ValueError("expected an nsamples x nparams array")
    n, npar = chain.shape
    if n > lib.pxm_quantile_columns_max_samples():
        return _host_range(chain, qa, qb)
    D.dev()
    out = np.empty(npar, dtype=np.float64)
    step = max(32, int(max_bytes // (8 * max(n, 1))) // 32 * 32)  # columns per upload
    for c0 in range(0, npar, step):
        blk = torch.from_numpy(np.ascontiguousarray(chain[:, c0:c0 + step], dtype=np.float64)).to(D.dev())
        lo, hi = _quantile_pair_dev(blk, qa, qb)
        out[c0:c0 + step] = D.to_host(hi - lo)
    return out


def wavelet_credible_interval_range(chain, L, B, J_min, alpha=0.05):
    """Credible-interval range of every wavelet coefficient, one MW (theta, phi) map per scale
    (pxmcmc/uncertainty.py:19-42)."""
    rng = credible_interval_range(chain, alpha)
    maps, start = [], 0
    for bl in _multires_bandlimits(L, B, J_min):
        length = bl * (2 * bl - 1)
        maps.append(rng[start:start + length].reshape((bl, 2 * bl - 1)))
        start += length
    return maps


def credible_region_threshold(logpis, alpha=0.05):
    """log-posterior threshold of the (1 - alpha) credible set (pxmcmc/uncertainty.py:45-53)"""
    return np.quantile(np.asarray(logpis), 1 - alpha)


def in_credible_region(logpi, threshold):
    """:meta private:"""
    return True if logpi <= threshold else False


def chain_to_pixels(chain, transform, batch=64):
    """Apply ``transform.inverse`` (wavelet synthesis) to every saved sample, ``batch`` samples per
    launch, and keep the real part as the reference's plot scripts do when they assign into a float
    array (experiments/earthtopography/plot.py:103-109).  Returns nsamples x npix float64 (numpy)."""
    chain = np.asarray(chain)
    n = chain.shape[0]
    out = None
    for i0 in range(0, n, batch):
        x = D.to_dev_c(chain[i0:i0 + batch])
        pix = D.to_host(transform.inverse(x).real)
        if out is None:
            out = np.empty((n, pix.shape[-1]), dtype=np.float64)
        out[i0:i0 + batch] = pix
    return out
