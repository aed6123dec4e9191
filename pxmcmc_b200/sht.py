"""pyssht-level spin spherical harmonic transforms on MW sampling, GPU-resident.

Mirrors the four pyssht calls of the reference's hot path
(pxmcmc/measurements.py:223,225,237,239): ``forward``, ``inverse``,
``inverse_adjoint``, ``forward_adjoint`` with ``Method="MW"``.  numpy in ->
numpy out; CUDA tensors ([..] or [nbatch, ..]) stay on the device."""
import numpy as np

from . import device as D


def sample_length(L, Method="MW"):
    return L * (2 * L - 1)


def sample_shape(L, Method="MW"):
    return (L, 2 * L - 1)


def sample_positions(L, Method="MW", Grid=False):
    n = 2 * L - 1
    thetas = (2.0 * np.arange(L) + 1.0) * np.pi / n
    phis = 2.0 * np.pi * np.arange(n) / n
    if Grid:
        return np.meshgrid(thetas, phis, indexing="ij")
    return thetas, phis


def elm2ind(el, m):
    return el * el + el + m


def _check(Method):
    if Method != "MW":
        raise NotImplementedError("only MW sampling is implemented")


def forward(f, L, Spin=0, Method="MW", Reality=False):
    _check(Method)
    x = D.to_dev_c(f)
    flat = x.reshape(-1) if x.numel() == L * (2 * L - 1) else x.reshape(-1, L * (2 * L - 1))
    nb = 1 if flat.dim() == 1 else flat.shape[0]
    return D.like_input(D.ShtPlan.get(L, Spin, nb).forward(flat), f)


def inverse(flm, L, Spin=0, Method="MW", Reality=False):
    _check(Method)
    x = D.to_dev_c(flm)
    nb = 1 if x.dim() == 1 else x.shape[0]
    out = D.ShtPlan.get(L, Spin, nb).inverse(x)
    if D.is_dev(flm):
        return out
    res = D.to_host(out).reshape(L, 2 * L - 1)
    return res.real.copy() if Reality else res


def inverse_adjoint(f, L, Spin=0, Method="MW", Reality=False):
    _check(Method)
    x = D.to_dev_c(f)
    flat = x.reshape(-1) if x.numel() == L * (2 * L - 1) else x.reshape(-1, L * (2 * L - 1))
    nb = 1 if flat.dim() == 1 else flat.shape[0]
    return D.like_input(D.ShtPlan.get(L, Spin, nb).inverse_adjoint(flat), f)


def forward_adjoint(flm, L, Spin=0, Method="MW", Reality=False):
    _check(Method)
    x = D.to_dev_c(flm)
    nb = 1 if x.dim() == 1 else x.shape[0]
    out = D.ShtPlan.get(L, Spin, nb).forward_adjoint(x)
    if D.is_dev(flm):
        return out
    res = D.to_host(out).reshape(L, 2 * L - 1)
    return res.real.copy() if Reality else res
