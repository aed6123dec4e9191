"""Stand-in module named ``pyssht`` so that the reference's unmodified
``pxmcmc/*.py`` can be imported in this container (the real wheel is absent).
Delegates to the oracle restatement.  TEST INFRASTRUCTURE ONLY; used by
``oracle/gen_golden.py`` and ``oracle/refloader.py``."""
import numpy as np

from oracle import ssht_ref as _s


def sample_length(L, Method="MW"):
    return _s.sample_length(L)


def sample_shape(L, Method="MW"):
    return _s.sample_shape(L)


def sample_positions(L, Method="MW", Grid=False):
    thetas, phis = _s.sample_positions(L)
    if Grid:
        return np.meshgrid(thetas, phis, indexing="ij")
    return thetas, phis


def elm2ind(el, m):
    return _s.elm2ind(el, m)


def ind2elm(ind):
    return _s.ind2elm(ind)


def theta_to_index(theta, L, Method="MW"):
    return _s.theta_to_index(theta, L)


def phi_to_index(phi, L, Method="MW"):
    return _s.phi_to_index(phi, L)


def forward(f, L, Spin=0, Method="MW", Reality=False):
    assert Method == "MW"
    return _s.forward(f, L, Spin)


def inverse(flm, L, Spin=0, Method="MW", Reality=False):
    assert Method == "MW"
    f = _s.inverse(flm, L, Spin)
    return f.real.copy() if Reality else f


def inverse_adjoint(f, L, Spin=0, Method="MW", Reality=False):
    assert Method == "MW"
    return _s.inverse_adjoint(f, L, Spin)


def forward_adjoint(flm, L, Spin=0, Method="MW", Reality=False):
    assert Method == "MW"
    f = _s.forward_adjoint(flm, L, Spin)
    return f.real.copy() if Reality else f
