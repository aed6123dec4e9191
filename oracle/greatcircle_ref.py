"""numpy restatement of the great-circle path rasteriser behind
``experiments/phasevel/main.py:40-59`` of the reference.  TEST INFRASTRUCTURE
ONLY (see oracle/__init__.py).

The reference builds every row of the path matrix with
``greatcirclepaths.GreatCirclePath(start, stop, "MW", L=L, weighting="average",
latlon=True)``, ``get_points(points_per_rad=160)``, ``fill()`` and stacks the
``.map`` arrays into a scipy CSR (``phasevel/main.py:44-59``).
``greatcirclepaths==1.1.0`` (``poetry.lock:557-558``) is not in the image and its
source is not under /root/reference => against that wheel this file is *parity
unpinned*; it restates the published behaviour: points spaced evenly along the
minor arc at ``points_per_rad`` samples per radian of epicentral distance, each
binned to its nearest MW pixel (``pyssht.theta_to_index`` / ``phi_to_index``),
``weighting="average"`` = every pixel weighted by its share of the points, so a
row sums to one and ``A @ x`` is the path average of ``x``.  It is pinned by the
property the reference's own test asserts for path matrices (a constant map
integrates to the path length / averages to the constant,
``tests/test_measurements.py:32-45``) and by geometry (all points at the
endpoints' distance from the pole of the circle).
"""
import numpy as np
from scipy import sparse


def _unit(lat_deg, lon_deg):
    lat, lon = np.radians(lat_deg), np.radians(lon_deg)
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)


def path_points(start, stop, points_per_rad=160):
    """start, stop: (lat, lon) in degrees.  Returns (theta, phi) of
    n = max(2, ceil(points_per_rad * distance)) points evenly spaced on the
    minor arc, end points included (spherical linear interpolation)."""
    a, b = _unit(*start), _unit(*stop)
    d = np.arccos(np.clip(np.dot(a, b), -1.0, 1.0))
    n = max(2, int(np.ceil(points_per_rad * d)))
    f = np.arange(n) / (n - 1.0)
    if d < 1e-12:
        p = np.repeat(a[None, :], n, axis=0)
    else:
        p = (np.sin((1 - f) * d)[:, None] * a[None, :] + np.sin(f * d)[:, None] * b[None, :]) / np.sin(d)
    theta = np.arccos(np.clip(p[:, 2], -1.0, 1.0))
    phi = np.mod(np.arctan2(p[:, 1], p[:, 0]), 2 * np.pi)
    return theta, phi


def pixel_indices(theta, phi, L):
    """nearest MW pixel of every point, flat index t*(2L-1)+p (pyssht.theta_to_index / phi_to_index)"""
    n = 2 * L - 1
    t = np.floor((theta * n / np.pi - 1.0) / 2.0 + 0.5).astype(np.int64)
    t = np.clip(t, 0, L - 1)
    p = np.floor(phi * n / (2 * np.pi) + 0.5).astype(np.int64) % n
    return t * n + p


def path_row(start, stop, L, points_per_rad=160):
    """(columns sorted, weights) of one row, weighting="average": share of the path's points per pixel"""
    theta, phi = path_points(start, stop, points_per_rad)
    pix = pixel_indices(theta, phi, L)
    cols, counts = np.unique(pix, return_counts=True)
    return cols, counts / float(pix.size)


def path_matrix(starts, stops, L, points_per_rad=160):
    """scipy CSR (npaths x L(2L-1)), the matrix of phasevel/main.py:50-59"""
    indptr, indices, data = [0], [], []
    for s, e in zip(starts, stops):
        c, w = path_row(s, e, L, points_per_rad)
        indices.append(c)
        data.append(w)
        indptr.append(indptr[-1] + c.size)
    return sparse.csr_matrix((np.concatenate(data), np.concatenate(indices), np.asarray(indptr)),
                             shape=(len(indptr) - 1, L * (2 * L - 1)))


def random_endpoints(npaths, seed=7):
    """uniformly random end-point pairs on the sphere as (lat, lon) degrees"""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1, 1, size=(npaths, 2))
    lon = rng.uniform(-180, 180, size=(npaths, 2))
    lat = np.degrees(np.arcsin(z))
    return np.stack([lat[:, 0], lon[:, 0]], axis=1), np.stack([lat[:, 1], lon[:, 1]], axis=1)
