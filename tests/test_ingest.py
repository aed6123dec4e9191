"""Data ingestion: FITS reader / writer, NESTED <-> RING, the Euclid-like mask and the great-circle
oracle on the CPU; the great-circle rasteriser kernel, hp.smoothing and the drivers' data preparation on the GPU."""
import os

import numpy as np
import pytest

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ETOPO = os.path.join(ROOT, "tests", "data", "ETOPO1_Ice_hpx_256.fits")  # the reference's own input (experiments/earthtopography/)


# ------------------------------------------------------------------ CPU
def test_reads_the_reference_topography_file():
    from pxmcmc_b200 import ingest

    m, hdr = ingest.read_map(ETOPO, h=True)
    hdr = dict(hdr)
    assert m.shape == (12 * 256 ** 2,) and m.dtype == np.float64
    assert hdr["NSIDE"] == 256 and hdr["ORDERING"] == "RING" and hdr["TFORM1"] == "1024E"
    # ETOPO1 "ice surface": deepest trench, highest summit at this resolution, mean elevation of the Earth's surface
    assert -11000 < m.min() < -9000 and 5500 < m.max() < 8900 and -2500 < m.mean() < -2300
    # big-endian float32 table rows, 1024 pixels per row: first pixels are the north-polar ice cap / Arctic ocean
    raw = np.frombuffer(open(ETOPO, "rb").read(), dtype=">f4", count=4, offset=2 * 2880)
    assert np.array_equal(m[:4], raw.astype(float))


def test_fits_write_read_round_trip_ring_and_nested(tmp_path):
    from pxmcmc_b200 import ingest

    rng = np.random.default_rng(0)
    for nside in (1, 4, 32):
        m = rng.standard_normal(12 * nside * nside)
        f = str(tmp_path / f"m{nside}.fits")
        ingest.write_map(f, m, dtype=np.float64)
        assert np.array_equal(ingest.read_map(f), m)
        assert os.path.getsize(f) % 2880 == 0
        order = ingest.nest2ring_order(nside)
        assert sorted(order) == list(range(m.size))
        ingest.write_map(f, m[order], nest=True, dtype=np.float64)      # the same sky stored in NESTED order
        assert np.array_equal(ingest.read_map(f), m)                    # returned in RING order
        assert np.array_equal(ingest.read_map(f, nest=True), m[order])
    ingest.write_map(str(tmp_path / "f32.fits"), np.arange(12.0) + 0.1)
    assert np.array_equal(ingest.read_map(str(tmp_path / "f32.fits")), (np.arange(12.0) + 0.1).astype(np.float32).astype(float))


def test_nested_order_is_hierarchical():
    """defining property of the NESTED scheme: pixel p at nside/2 is made of pixels 4p .. 4p+3 at nside -- checked with
    the RING pixel centres of the HEALPix primer (an independent formula)"""
    from oracle import healpix_ref
    from pxmcmc_b200 import ingest

    def vec(nside):
        th, ph = healpix_ref.pix2ang(nside)
        return np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])

    for nside in (2, 4, 8, 16):
        child = vec(nside)[:, ingest.nest2ring_order(nside)]
        parent = vec(nside // 2)[:, ingest.nest2ring_order(nside // 2)][:, np.arange(12 * nside * nside) // 4]
        ang = np.arccos(np.clip((child * parent).sum(0), -1, 1))
        assert ang.max() < 0.55 * np.sqrt(4 * np.pi / (12 * (nside // 2) ** 2))


def test_mask_and_galactic_rotation():
    from pxmcmc_b200 import ingest, utils

    M = ingest.icrs_to_galactic_matrix()
    assert np.abs(M @ M.T - np.eye(3)).max() < 1e-15 and abs(np.linalg.det(M) - 1) < 1e-15

    def u(ra, dec):
        ra, dec = np.radians(ra), np.radians(dec)
        return np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])

    # Sgr A* (ICRS 266.41683, -29.00781) lies at l = 359.944, b = -0.046; the north galactic pole at b = 90
    g = M @ u(266.41683, -29.00781)
    assert abs(np.degrees(np.arcsin(g[2])) + 0.046) < 2e-3 and abs(np.degrees(np.arctan2(g[1], g[0])) % 360 - 359.944) < 2e-3
    assert (M @ u(192.85948, 27.12825))[2] > 1 - 1e-12
    L = 48
    mask = utils.build_mask(L, 10)
    assert mask.shape == (L, 2 * L - 1) and set(np.unique(mask)) == {0.0, 1.0}
    th = np.degrees(utils.mw_sample_positions(L)[0])
    assert np.all(mask[np.abs(90 - th) < 10] == 0)                      # the "ecliptic" band of the reference's mask
    assert 0.6 < mask.mean() < 0.8                                      # two 20-degree great-circle bands
    assert mask[0].min() == 1 or mask[-1].min() == 1                    # the poles of the grid are not both masked
    # 20-degree half-width masks more
    assert utils.build_mask(L, 20).sum() < mask.sum()


def test_beam_and_almxfl():
    from pxmcmc_b200 import ingest, utils

    lmax = 5
    b = ingest.gauss_beam(np.radians(1.0), lmax)
    assert b[0] == 1.0 and np.allclose(b[3], np.exp(-0.5 * 12 * np.radians(1.0) ** 2))
    alm = np.arange(utils.alm_hp_size(lmax)) + 1j
    out = ingest.almxfl(alm, np.arange(lmax + 1.0), lmax)
    for m in range(lmax + 1):
        for el in range(m, lmax + 1):
            assert out[utils.alm_hp_index(el, m, lmax)] == alm[utils.alm_hp_index(el, m, lmax)] * el


def test_great_circle_oracle_properties():
    """pins oracle/greatcircle_ref.py: rows sum to one, a constant map averages to the constant (the property of
    tests/test_measurements.py:32-45 in "average" weighting), points lie on the great circle, meridian paths stay in
    one or two phi columns"""
    from oracle import greatcircle_ref as G

    L = 24
    st, sp = G.random_endpoints(50, seed=1)
    A = G.path_matrix(st, sp, L)
    assert np.allclose(np.asarray(A.sum(axis=1)).ravel(), 1.0, atol=1e-14)
    assert np.allclose(A @ np.full(L * (2 * L - 1), 3.5), 3.5)
    th, ph = G.path_points(st[0], sp[0])
    p = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    normal = np.cross(p[0], p[-1])
    assert np.abs(p @ normal).max() < 1e-12 and len(th) == max(2, int(np.ceil(160 * np.arccos(np.clip(p[0] @ p[-1], -1, 1)))))
    cols, w = G.path_row((80.0, 30.0), (-80.0, 30.0), L)  # along a meridian
    assert len(set(cols % (2 * L - 1))) <= 2 and len(set(cols // (2 * L - 1))) >= L - 4


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("L,npaths", [(28, 300), (128, 2000)])
def test_great_circle_rasteriser_matches_the_oracle(gpu, L, npaths):
    """pxm_gc_rasterise (one CTA per path) against the numpy rasteriser: the CSR of experiments/phasevel/main.py:40-59"""
    from oracle import greatcircle_ref as G
    from pxmcmc_b200 import paths

    st, sp = G.random_endpoints(npaths, seed=7)
    # edge cases: zero-length path, nearly antipodal end points, over the north pole, across the phi = 0 seam
    # (exactly antipodal end points define no great circle, and a path along the phi = pi meridian sits on a pixel
    # boundary of the odd-length rings: both are rounding coin flips in any implementation, so they are not test cases)
    st[:5] = [(10.0, 20.0), (0.0, 0.0), (80.0, 10.0), (-5.0, 359.0), (45.0, -170.0)]
    sp[:5] = [(10.0, 20.0), (0.5, 179.0), (70.0, 190.5), (5.0, 1.0), (44.0, 170.0)]
    A = paths.get_path_matrix(st, sp, L)
    R = G.path_matrix(st, sp, L)
    assert A.shape == R.shape == (npaths, L * (2 * L - 1))
    n_ours = paths.path_points_count(st, sp)
    n_ref = np.array([len(G.path_points(a, b)[0]) for a, b in zip(st, sp)])
    assert np.array_equal(n_ours, n_ref)
    assert np.allclose(np.asarray(A.sum(axis=1)).ravel(), 1.0, atol=1e-14)
    # identical sparsity pattern and weights (a point within an ulp of a pixel boundary may fall on either side:
    # allow one such point per ten thousand)
    D = (A - R).tocsr()
    D.eliminate_zeros()
    bad_rows = np.unique(D.nonzero()[0])
    assert bad_rows.size <= max(1, npaths // 1000), f"{bad_rows.size} rows differ"
    assert abs(D).sum() <= 2.0 * bad_rows.size / 100
    assert A.has_sorted_indices and np.array_equal(A.indptr[:6], R.indptr[:6])
    assert paths.get_path_matrix(np.zeros((0, 2)), np.zeros((0, 2)), L).shape == (0, L * (2 * L - 1))


@pytest.mark.gpu
def test_phasevel_data_preparation_end_to_end(gpu, tmp_path):
    """experiments/phasevel/main.py:22-59,128-140: datafile -> path matrix -> PathIntegralOperator predictions"""
    from pxmcmc_b200 import paths
    from pxmcmc_b200.forward import PathIntegralOperator

    L, B, J = 28, 2, 2
    rng = np.random.default_rng(3)
    n = 200
    rows = np.column_stack([rng.uniform(-80, 80, n), rng.uniform(-180, 180, n), rng.uniform(-80, 80, n),
                            rng.uniform(-180, 180, n), rng.standard_normal(n), -np.abs(rng.standard_normal(n)) - 0.1,
                            np.ones(n), rng.integers(1, 5, n)])
    f = str(tmp_path / "paths.txt")
    np.savetxt(f, rows)
    with pytest.warns(UserWarning):
        start, stop, data, sig_d, mima, nsim = paths.read_datafile(f)
    assert np.all(sig_d > 0) and start.shape == (n, 2)
    A = paths.get_path_matrix(start, stop, L)
    op = PathIntegralOperator(A, data, sig_d, "synthesis", L, B, J)
    const = np.full(L * (2 * L - 1), 2.0)
    assert np.allclose(op.measurement.forward(const), 2.0)
    X = rng.standard_normal(op.nparams)
    assert rel_l2(op.forward(X), A @ op.transform.inverse(X)) < 1e-13


@pytest.mark.gpu
def test_smoothing_against_the_dense_oracle(gpu):
    """hp.smoothing = map2alm(iter=3) -> Gaussian beam -> alm2map (experiments/weaklensing/main.py:34-36)"""
    from oracle import healpix_ref
    from pxmcmc_b200 import ingest

    nside, lmax = 8, 15
    rng = np.random.default_rng(4)
    m = rng.standard_normal(12 * nside * nside)
    sigma = np.radians(5.0)
    ours = ingest.smoothing(m, sigma=sigma, lmax=lmax)
    Y = healpix_ref.ylm_matrix(nside, lmax + 1)
    alm = healpix_ref.map2alm(m, lmax, iter=3, Y=Y)
    ref = healpix_ref.alm2map(ingest.almxfl(alm, ingest.gauss_beam(sigma, lmax), lmax), nside, Y=Y)
    assert rel_l2(ours, ref) < 1e-10
    assert np.allclose(ingest.smoothing(m, fwhm=sigma * 2 * np.sqrt(2 * np.log(2)), lmax=lmax), ours, rtol=0, atol=1e-13)
    assert np.std(ours) < np.std(m)


@pytest.mark.gpu
def test_earthtopography_data_preparation_on_the_real_file(gpu):
    """experiments/earthtopography/main.py:79-82,119 on the reference's own ETOPO1 file: read_map -> map2alm(L - 1) ->
    lm_hp2lm -> alm2map_mw, the data of config 1 (L = 32) and of config 2 (L = 256)"""
    from pxmcmc_b200 import ingest, utils
    from pxmcmc_b200.forward import SphericalWaveletTransformOperator

    topo = ingest.read_map(ETOPO)
    for L in (32, 256):
        alm = utils.map2alm(topo, L - 1)
        flm = utils.lm_hp2lm(alm, L)
        topo_d = utils.alm2map_mw(flm, L, 0)
        assert topo_d.shape == (L * (2 * L - 1),) and np.iscomplexobj(topo_d)
        assert np.abs(topo_d.imag).max() < 1e-8 * np.abs(topo_d.real).max()          # a real field
        # the monopole is the mean elevation; the bandlimited map keeps the ocean / continent contrast
        assert abs(flm[0].real / np.sqrt(4 * np.pi) - topo.mean()) < 1.0
        assert topo_d.real.min() < -4000 and topo_d.real.max() > 2500
        # Jacobi-refined analysis: synthesising the a_lm back reproduces the bandlimited part of the input
        back = utils.alm2map(alm, 256)
        again = utils.map2alm(back, L - 1)
        assert rel_l2(again, alm) < 1e-6
        op = SphericalWaveletTransformOperator(topo_d / 1000, 1.0, "synthesis", L, 1.5, 2)
        assert np.iscomplexobj(op.invcov.diagonal())                                    # forward.py:80-82: complex data, real sigma
    # Himalaya / Tibet (theta ~ 57 deg, phi ~ 88 deg) is high, the mid-Pacific is deep (L = 256 map)
    th, ph = utils.mw_sample_positions(256)
    f = topo_d.real.reshape(256, 511)
    assert f[np.argmin(abs(th - np.radians(57))), np.argmin(abs(ph - np.radians(88)))] > 3000
    assert f[np.argmin(abs(th - np.radians(90))), np.argmin(abs(ph - np.radians(200)))] < -3000


@pytest.mark.gpu
def test_weaklensing_data_preparation(gpu, tmp_path):
    """experiments/weaklensing/main.py:23-39,88-90 with a synthetic HEALPix convergence file and build_mask"""
    from pxmcmc_b200 import ingest, utils
    from pxmcmc_b200.measurements import WeakLensing

    L, nside = 32, 32
    rng = np.random.default_rng(6)
    flm = np.zeros(L * L, dtype=complex)
    for el in range(2, L):
        flm[el * el + el] = rng.standard_normal()
        m = np.arange(1, el + 1)
        a = (rng.standard_normal(el) + 1j * rng.standard_normal(el)) / np.sqrt(2)
        flm[el * el + el + m] = a
        flm[el * el + el - m] = (-1.0) ** m * np.conj(a)
    kappa_hp = utils.alm2map(utils.lm2lm_hp(flm, L), nside)
    f = str(tmp_path / "kappa.fits")
    ingest.write_map(f, kappa_hp, dtype=np.float64)
    mask = utils.build_mask(L, size=10)
    wl = WeakLensing(L, mask, ngal=np.full_like(mask, 30))
    gam = ingest.load_gammas(f, L, wl)
    assert gam.shape == (int(mask.sum()),) and np.iscomplexobj(gam)
    # the same through explicit steps: beam on the known flm (map2alm of a bandlimited map at nside = lmax + 1 is
    # accurate to the Jacobi residual), MW synthesis, operator
    els = np.floor(np.sqrt(np.arange(L * L))).astype(int)
    ref = wl.forward(utils.alm2map_mw(flm * ingest.gauss_beam(np.radians(50 / 60), L - 1)[els], L))
    assert rel_l2(gam, ref) < 1e-3
