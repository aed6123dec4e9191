"""Great-circle path matrices for ``PathIntegral`` -- the data preparation of the reference's phase-velocity
experiment (``experiments/phasevel/main.py:22-59``) without the ``greatcirclepaths`` wheel and its worker pool.

``get_path_matrix(start, stop, L)`` has the signature and the result (a scipy CSR, one row per path, MW pixels as
columns, rows summing to one) of the reference's function of that name; every path is rasterised by one CTA of
``pxm_gc_rasterise`` (libpxmcmc_b200.so)."""
from warnings import warn

import numpy as np
import torch
from scipy import sparse

from . import device as D
from ._lib import check, lib, ptr, stream_ptr


def read_datafile(datafile):
    """columns start_lat, start_lon, stop_lat, stop_lon, data, error, minor/major, n_similar; degrees
    (experiments/phasevel/main.py:22-37)"""
    start_lat, start_lon, stop_lat, stop_lon, data, sig_d, mima, nsim = np.loadtxt(datafile, unpack=True)
    start = np.stack([start_lat, start_lon], axis=1)
    stop = np.stack([stop_lat, stop_lon], axis=1)
    if np.any(sig_d < 0):
        warn("Some of the data errors read in are negative. Forcing positivity.")
        sig_d = np.abs(sig_d)
    return start, stop, data, sig_d, mima, nsim


def get_path_matrix(start, stop, L=32, processes=None, points_per_rad=160, device_csr=False):
    """Matrix of all the great-circle paths (the measurement operator of ``PathIntegral``):
    row r = the MW pixels path r passes through, weighted by the share of the path's points that fall into each
    (``weighting="average"``), points every 1/points_per_rad radians (experiments/phasevel/main.py:40-59).

    :param start, stop: npaths x 2 arrays of (latitude, longitude) in degrees
    :param processes: ignored (the reference's multiprocessing pool size)
    :param device_csr: also return the CSR arrays as CUDA tensors (indptr int64, indices int32, data float64)
    """
    start = np.ascontiguousarray(np.asarray(start, dtype=np.float64).reshape(-1, 2))
    stop = np.ascontiguousarray(np.asarray(stop, dtype=np.float64).reshape(-1, 2))
    if start.shape != stop.shape:
        raise ValueError("start and stop must hold one (lat, lon) pair per path")
    npaths = start.shape[0]
    npix = L * (2 * L - 1)
    dv = D.dev()
    if npaths == 0:
        return sparse.csr_matrix((0, npix))
    cap = 1 << int(np.ceil(np.log2(max(2.0, np.ceil(points_per_rad * np.pi) + 1))))
    if cap > 4096:
        raise ValueError("points_per_rad too large for the rasteriser (at most 1303 points per radian)")
    s_d, e_d = torch.from_numpy(start).to(dv), torch.from_numpy(stop).to(dv)
    cols = torch.empty((npaths, cap), dtype=torch.int32, device=dv)
    w = torch.empty((npaths, cap), dtype=torch.float64, device=dv)
    nnz = torch.empty(npaths, dtype=torch.int32, device=dv)
    check(lib.pxm_gc_rasterise(ptr(s_d), ptr(e_d), npaths, int(L), float(points_per_rad), cap, ptr(cols), ptr(w), ptr(nnz),
                               stream_ptr()))
    indptr = np.zeros(npaths + 1, dtype=np.int64)
    np.cumsum(nnz.cpu().numpy(), out=indptr[1:])  # set-up time: the row pointer is formed on the host
    ip_d = torch.from_numpy(indptr).to(dv)
    indices = torch.empty(int(indptr[-1]), dtype=torch.int32, device=dv)
    data = torch.empty(int(indptr[-1]), dtype=torch.float64, device=dv)
    check(lib.pxm_gc_compact(ptr(ip_d), ptr(cols), ptr(w), cap, npaths, ptr(indices), ptr(data), stream_ptr()))
    A = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr), shape=(npaths, npix))
    A.has_sorted_indices = True
    return (A, (ip_d, indices, data)) if device_csr else A


def path_points_count(start, stop, points_per_rad=160):
    """number of points ``get_points(points_per_rad)`` places on every path"""
    start = np.ascontiguousarray(np.asarray(start, dtype=np.float64).reshape(-1, 2))
    stop = np.ascontiguousarray(np.asarray(stop, dtype=np.float64).reshape(-1, 2))
    dv = D.dev()
    out = torch.empty(start.shape[0], dtype=torch.int32, device=dv)
    s_d, e_d = torch.from_numpy(start).to(dv), torch.from_numpy(stop).to(dv)  # named: they must outlive the launch
    check(lib.pxm_gc_count_points(ptr(s_d), ptr(e_d), start.shape[0], float(points_per_rad), ptr(out), stream_ptr()))
    return out.cpu().numpy()
