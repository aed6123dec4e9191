"""CPU: the oracle restatement (oracle/pxmcmc_ref.py) against fixtures produced
by the UNMODIFIED reference (oracle/gen_golden.py), plus the reference's own
known-answer / property tests re-expressed (tests/test_utils.py,
test_transforms.py, test_measurements.py of the reference)."""
import numpy as np
import pytest
from scipy import sparse

from conftest import golden, rel_l2
from oracle import pxmcmc_ref as R
from oracle import s2let_ref, ssht_ref

TOL = 1e-12


def test_soft_known_answers():
    # reference tests/test_utils.py:35-44 (exact ==)
    assert all(R.soft([1, 2, 3], 2) == [0, 0, 1])
    assert all(R.soft([-1, -2, -3], 2) == [0, 0, -1])
    assert all(R.soft([1 + 1j, 0.5 - 0.5j, 0], 1) == [(1 + 1j) * (np.sqrt(2) - 1) / np.sqrt(2), 0, 0])


def test_soft_golden():
    g = golden("ref_soft.npz")
    assert np.array_equal(R.soft(g["xr"], 0.8), g["soft_r_scalar"])
    assert np.array_equal(R.soft(g["xr"], g["tv"]), g["soft_r_vec"])
    assert np.array_equal(R.soft(g["xc"], 1.0), g["soft_c_scalar"])
    assert np.array_equal(R.soft(g["xc"], g["tv"]), g["soft_c_vec"])


def test_chebyshev_known_answers():
    # reference tests/test_utils.py:54-66
    assert R.chebyshev1(5, 0) == 1 and R.chebyshev1(2, 1) == 2 and R.chebyshev1(3, 5) == 3363
    assert R.chebyshev2(5, 0) == 1 and R.chebyshev2(2, 1) == 4 and R.chebyshev2(3, 5) == 6930
    assert R.cheb1der(5, 0) == 0 and R.cheb1der(2, 1) == 1 and R.cheb1der(3, 5) == 5945


def test_quadrature_weights_integrate_to_4pi():
    for L in (4, 10, 33):
        assert np.isclose(R.mw_map_weights(L).sum(), 4 * np.pi, rtol=0, atol=1e-12)


def test_s2_integrate_property(rng):
    # reference tests/test_utils.py:85-100: pins SHT normalisation + quadrature jointly
    L = 10
    flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    f = ssht_ref.inverse(flm, L, 0).ravel()
    assert np.isclose((R.mw_map_weights(L) * f).sum(), flm[0] * np.sqrt(4 * np.pi))


@pytest.mark.parametrize("spin", [0, 2, -2])
def test_sht_against_closed_form_and_properties(rng, spin):
    L = 7
    flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    flm[: spin * spin] = 0
    f = ssht_ref.inverse(flm, L, spin)
    assert np.abs(f - ssht_ref.inverse_direct(flm, L, spin)).max() < 1e-12
    assert np.abs(ssht_ref.forward(f, L, spin) - flm).max() < 1e-12
    g = rng.standard_normal(f.shape) + 1j * rng.standard_normal(f.shape)
    h = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    h[: spin * spin] = 0
    assert abs(np.vdot(g, f) - np.vdot(ssht_ref.inverse_adjoint(g, L, spin), flm)) < 1e-11
    assert abs(np.vdot(h, ssht_ref.forward(g, L, spin)) - np.vdot(ssht_ref.forward_adjoint(h, L, spin), g)) < 1e-11


def test_sht_forward_equals_quadrature_operator(rng):
    """`forward` on NON-bandlimited input equals the dense per-m operator built
    from fine-grid quadrature (the construction the GPU tables use)."""
    L, spin = 6, 2
    n = 2 * L - 1
    g = rng.standard_normal((L, n)) + 1j * rng.standard_normal((L, n))
    ms = np.arange(-(L - 1), L)
    Fm = g @ np.exp(-1j * np.outer(2 * np.pi * np.arange(n) / n, ms)) / n
    flm = ssht_ref.forward(g, L, spin)
    for m in ms:
        out = ssht_ref.forward_quadrature_matrix(L, spin, int(m)) @ Fm[:, m + L - 1]
        for l in range(max(abs(m), abs(spin)), L):
            assert abs(out[l] - flm[l * l + l + m]) < 1e-12


def test_sht_golden():
    g = golden("ref_sht_L12.npz")
    L = int(g["L"])
    for s in (0, 2):
        assert rel_l2(ssht_ref.inverse(g[f"flm_s{s}"], L, s), g[f"inverse_s{s}"]) < TOL
        assert rel_l2(ssht_ref.forward(g["f"], L, s), g[f"forward_s{s}"]) < TOL
        assert rel_l2(ssht_ref.inverse_adjoint(g["f"], L, s), g[f"inverse_adjoint_s{s}"]) < TOL
        assert rel_l2(ssht_ref.forward_adjoint(g[f"flm_s{s}"], L, s), g[f"forward_adjoint_s{s}"]) < TOL


@pytest.mark.parametrize("LBJ", [(10, 2, 2), (32, 1.5, 2), (128, 2, 2), (256, 1.5, 2)])
def test_tiling_partition_of_unity_and_bandlimits(LBJ):
    L, B, J = LBJ
    kap, k0 = s2let_ref.tiling_axisym(B, L, J)
    assert np.abs(k0 ** 2 + (kap ** 2).sum(0) - 1).max() < 1e-14
    expect = {
        (10, 2, 2): [4, 8, 10, 10],
        (32, 1.5, 2): [3, 4, 6, 8, 12, 18, 26, 32, 32],
        (128, 2, 2): [4, 8, 16, 32, 64, 128, 128],
        (256, 1.5, 2): [3, 4, 6, 8, 12, 18, 26, 39, 58, 87, 130, 195, 256, 256],
    }[LBJ]
    assert s2let_ref.bandlimits(B, L, J) == expect
    # support-derived bandlimits (reference utils._multires_bandlimits) agree
    sup = [int(np.nonzero(k0)[0].max()) + 1] + [int(np.nonzero(k)[0].max()) + 1 for k in kap[J:]]
    assert sup == expect


@pytest.mark.parametrize("tag", ["L10B2", "L16B1p5"])
def test_wavelet_golden_and_properties(tag, rng):
    g = golden(f"ref_wavelet_{tag}.npz")
    L, B, J = int(g["L"]), float(g["B"]), int(g["J_min"])
    t = R.WaveletTransform(L, B, J)
    assert (t.nscal, t.nwav) == (int(g["nscal"]), int(g["nwav"]))
    assert rel_l2(t.forward(g["x_pix"]), g["forward"]) < TOL
    assert rel_l2(t.inverse(g["x_coef"]), g["inverse"]) < TOL
    assert rel_l2(t.inverse_adjoint(g["x_pix"]), g["inverse_adjoint"]) < TOL
    assert rel_l2(t.forward_adjoint(g["x_coef"]), g["forward_adjoint"]) < TOL
    # reference tests/test_transforms.py:16-46
    x = s2let_ref.alm2map_mw(ssht_ref.forward(g["x_pix"].reshape(L, -1), L, 0), L, 0)
    assert rel_l2(t.inverse(t.forward(x)), x) < 1e-11
    f = g["x_coef"]
    assert abs(np.vdot(f, t.forward(x)) - np.vdot(t.forward_adjoint(f), x)) < 1e-9
    assert abs(np.vdot(x, t.inverse(f)) - np.vdot(t.inverse_adjoint(x), f)) < 1e-9
    # prior weights
    assert rel_l2(R.S2WaveletsL1("synthesis", None, None, 1.0, L, B, J).T, g["s2_T"]) < 1e-14
    pw = R.S2WaveletsL1PowerWeights("synthesis", None, None, 1.0, L, B, J, eta=1)
    assert rel_l2(pw.T, g["s2pw_T"]) < 1e-14
    assert rel_l2(pw.map_weights, g["s2pw_w"]) < 1e-14
    assert rel_l2(R.mw_map_weights(L), g["mw_weights"]) < 1e-14


def test_flatten_layout_known_answer():
    # reference tests/test_utils.py:8-16: scaling first, then wavelets by scale
    g = golden("ref_flatten.npz")
    assert np.array_equal(g["flat"], np.concatenate([[i] * 861 for i in range(10)]))


def _myula_replay(g):
    """Re-run the chain with the oracle, drawing host noise in the reference's order."""
    from scipy.stats import laplace

    L, B, J = int(g["L"]), float(g["B"]), int(g["J_min"])
    t = R.WaveletTransform(L, B, J)
    sig = g["sig_d"] if g["sig_d"].ndim else float(g["sig_d"])
    op = R.ForwardOperator(g["data"], sig, "synthesis", t, R.IdentityMeasurement(len(g["data"]), L * (2 * L - 1)), t.ncoefs)
    delta, lmda, mu = float(g["delta"]), float(g["lmda"]), float(g["mu"])
    prior = R.S2WaveletsL1("synthesis", t.inverse, t.inverse_adjoint, lmda * mu, L, B, J)
    np.random.seed(int(g["seed"]))
    X = laplace.rvs(size=op.nparams)
    preds = op.forward(X)
    nsamples, nburn, ngap = int(g["nsamples"]), int(g["nburn"]), int(g["ngap"])
    i = j = 0
    out = {"chain": [], "logPi": [], "L2s": [], "priors": [], "preds": []}
    while j < nsamples:
        w = np.random.randn(op.nparams)
        X, preds = R.myula_iteration(op, prior, delta, lmda, X, preds, w)
        if i >= nburn and (ngap == 0 or (i - nburn) % ngap == 0):
            lp, l2, pr = R.logpi(op, prior, mu, X, preds)
            for k, v in zip(out, (X, lp, l2, pr, preds)):
                out[k].append(v)
            j += 1
        i += 1
    return {k: np.array(v) for k, v in out.items()}


@pytest.mark.parametrize("tag", ["L10_complex", "L10_real_sigvec", "L16B1p5_complex"])
def test_myula_chain_golden(tag):
    g = golden(f"ref_myula_{tag}.npz")
    out = _myula_replay(g)
    # the reference stores complex values into float arrays (real part kept)
    assert rel_l2(out["chain"].real, g["chain"]) < 1e-11
    assert rel_l2(out["preds"].real, g["preds"]) < 1e-11
    assert rel_l2(out["logPi"].real, g["logPi"]) < 1e-11
    assert rel_l2(out["L2s"].real, g["L2s"]) < 1e-11
    assert rel_l2(out["priors"].real, g["priors"]) < 1e-11


def test_myula_single_step_golden():
    g = golden("ref_myula_step_L10.npz")
    L, B, J = int(g["L"]), float(g["B"]), int(g["J_min"])
    t = R.WaveletTransform(L, B, J)
    op = R.ForwardOperator(g["data"], float(g["sig_d"]), "synthesis", t, R.IdentityMeasurement(len(g["data"]), len(g["data"])), t.ncoefs)
    prior = R.S2WaveletsL1("synthesis", t.inverse, t.inverse_adjoint, float(g["lmda"]) * float(g["mu"]), L, B, J)
    assert rel_l2(op.invcov, g["invcov"]) < 1e-15
    assert rel_l2(prior.T, g["T"]) < 1e-14
    assert rel_l2(op.forward(g["X"]), g["preds"]) < TOL
    assert rel_l2(op.calc_gradg(g["preds"]), g["gradg"]) < TOL
    px = prior.proxf(g["X"])
    assert np.array_equal(px == 0, g["prox"] == 0)
    assert rel_l2(px, g["prox"]) < 1e-14
    assert rel_l2(R.myula_step(g["X"], g["prox"], g["gradg"], float(g["delta"]), float(g["lmda"]), g["w"]), g["Xn"]) < 1e-15
    lp, l2, pr = R.logpi(op, prior, float(g["mu"]), g["X"], g["preds"])
    assert np.isclose(lp, g["logpi"], rtol=1e-12) and np.isclose(l2, g["L2"], rtol=1e-12) and np.isclose(pr, g["prior"], rtol=1e-12)


def test_pxmala_golden():
    from scipy.stats import laplace

    g = golden("ref_pxmala_L10.npz")
    L, B, J = int(g["L"]), float(g["B"]), int(g["J_min"])
    t = R.WaveletTransform(L, B, J)
    npix = L * (2 * L - 1)
    op = R.ForwardOperator(g["data"], float(g["sig_d"]), "analysis", t, R.IdentityMeasurement(npix, npix), npix)
    delta, lmda, mu = float(g["delta"]), float(g["lmda"]), float(g["mu"])
    prior = R.L1("analysis", t.inverse, t.inverse_adjoint, lmda * mu)
    assert np.isclose(R.pxmala_logtransition(g["X1"], g["X2"], g["proxf"], g["gradg"], delta, lmda), g["logtrans"], rtol=1e-12)
    assert rel_l2(prior.proxf(g["X1"]), g["proxf"]) < 1e-11
    # replay (mcmc.py:218-275)
    np.random.seed(int(g["seed"]))
    X = laplace.rvs(size=npix)
    preds = op.forward(X)
    gradg, px = op.calc_gradg(preds), prior.proxf(X)
    lp, l2, pr = R.logpi(op, prior, mu, X, preds)
    acc, deltas, chain = [], [delta], []
    nsamples, nburn, ngap = int(g["nsamples"]), int(g["nburn"]), int(g["ngap"])
    i = j = 0
    while j < nsamples:
        w = np.random.randn(npix)
        Xp = R.myula_step(X, px, gradg, delta, lmda, w)
        pp = op.forward(Xp)
        gp, pxp = op.calc_gradg(pp), prior.proxf(Xp)
        tcp = R.pxmala_logtransition(X, Xp, px, gradg, delta, lmda)
        tpc = R.pxmala_logtransition(Xp, X, pxp, gp, delta, lmda)
        lpp, l2p, prp = R.logpi(op, prior, mu, Xp, pp)
        ok = np.log(np.random.rand()) < tpc + lpp - tcp - lp
        if ok:
            X, preds, gradg, px, lp, l2, pr = Xp, pp, gp, pxp, lpp, l2p, prp
        acc.append(int(ok))
        delta = R.pxmala_tune_delta(delta, lmda, acc[i], i)
        deltas.append(delta)
        if i >= nburn and (ngap == 0 or (i - nburn) % ngap == 0) and ok:
            chain.append(X)
            j += 1
        i += 1
    assert acc == list(g["acceptance_trace"])
    assert np.allclose(deltas, g["deltas_trace"], rtol=1e-14, atol=0)
    assert rel_l2(np.array(chain).real, g["chain"]) < 1e-10


def test_skrock_golden():
    g = golden("ref_skrock_L10.npz")
    for s, sfx in ((3, ""), (5, "5")):
        w0, w1, mus, nus, ks = R.skrock_coefs(s)
        assert np.isclose(w0, g["omega_0" if s == 3 else "omega5_0"], rtol=1e-15)
        assert np.isclose(w1, g["omega_1" if s == 3 else "omega5_1"], rtol=1e-14)
        assert np.allclose(mus, g["mus" + sfx], rtol=1e-13) and np.allclose(nus, g["nus" + sfx], rtol=1e-13)
        assert np.allclose(ks, g["ks" + sfx], rtol=1e-13)
    L, B, J = int(g["L"]), float(g["B"]), int(g["J_min"])
    A = sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=tuple(g["A_shape"]))
    t = R.WaveletTransform(L, B, J)
    op = R.ForwardOperator(g["data"], g["sig_d"], "synthesis", t, R.PathIntegral(A), t.ncoefs)
    prior = R.S2WaveletsL1PowerWeights("synthesis", t.inverse, t.inverse_adjoint, float(g["lmda"]) * float(g["mu"]), L, B, J, eta=1)
    assert rel_l2(prior.T, g["T"]) < 1e-14
    assert np.isclose(prior.prior(g["X"]), g["prior"], rtol=1e-13)
    assert rel_l2(op.forward(g["X"]), g["preds"]) < TOL
    assert rel_l2(op.calc_gradg(g["preds"]), g["gradg"]) < TOL
    assert rel_l2(R.gradlogpi(op, prior, float(g["lmda"]), g["X"]), g["gradlogpi"]) < TOL
    Xn = R.skrock_step(op, prior, float(g["delta"]), float(g["lmda"]), 3, g["X"], g["Z"])
    assert rel_l2(Xn, g["Xn"]) < 1e-10


def test_weaklensing_golden(rng):
    g = golden("ref_weaklensing_L12.npz")
    L = int(g["L"])
    wl = R.WeakLensing(L, mask=g["mask"], ngal=g["ngal"])
    assert np.array_equal(wl.kernel, g["kernel"] * (np.arange(L * L) >= 4))
    assert rel_l2(wl.inv_cov, g["inv_cov"]) < 1e-15
    assert rel_l2(wl.forward(g["kappa"]), g["forward"]) < TOL
    assert rel_l2(wl.adjoint(g["gamma"]), g["adjoint"]) < TOL
    wl0 = R.WeakLensing(L)
    assert rel_l2(wl0.forward(g["kappa"]), g["forward_nomask"]) < TOL
    assert rel_l2(wl0.adjoint(g["kappa"]), g["adjoint_nomask"]) < TOL
    # reference tests/test_measurements.py:95-130 (masked dot test)
    a = abs(np.vdot(g["kappa"], wl.adjoint(g["gamma"])))
    b = abs(np.vdot(g["gamma"], wl.forward(g["kappa"])))
    assert np.isclose(a, b)
    t = R.WaveletTransform(L, 2, 2)
    fo = R.ForwardOperator(g["gdata"], 1 / wl.inv_cov, "synthesis", t, wl, t.ncoefs)
    assert rel_l2(fo.invcov, g["op_invcov"]) < 1e-15
    assert rel_l2(fo.forward(g["X"]), g["op_forward"]) < TOL
    assert rel_l2(fo.calc_gradg(g["op_forward"]), g["op_gradg"]) < TOL


def test_pathintegral_properties(rng):
    # reference tests/test_measurements.py:8-45
    L = 10
    A = sparse.random(100, L * (2 * L - 1), density=0.05, random_state=1, format="csr")
    p = R.PathIntegral(A)
    x = rng.standard_normal(L * (2 * L - 1))
    y = rng.random(100)
    assert np.isclose(y.conj().dot(p.forward(x)) - p.adjoint(y).conj().dot(x), 0)
    ring = np.zeros((L, 2 * L - 1))
    ring[ssht_ref.theta_to_index(np.pi / 2, L), :] = 2 * np.pi / (2 * L - 1)
    assert np.isclose(R.PathIntegral(sparse.csr_matrix(ring.reshape(1, -1))).forward(np.ones(L * (2 * L - 1))), 2 * np.pi)


# ---------------------------------------------------------------- HEALPix oracle
def test_healpix_oracle_pixel_centres_known_answers():
    """closed-form RING pixel centres of the HEALPix primer (nside 1 and 2) and equal-area rings"""
    from oracle import healpix_ref as H

    th, ph = H.pix2ang(1)
    assert np.allclose(np.cos(th), [2 / 3] * 4 + [0] * 4 + [-2 / 3] * 4, atol=1e-15)
    assert np.allclose(ph / np.pi, [0.25, 0.75, 1.25, 1.75, 0, 0.5, 1, 1.5, 0.25, 0.75, 1.25, 1.75])
    th, ph = H.pix2ang(2)
    assert np.allclose(np.cos(th[:4]), 1 - 1 / 12) and np.allclose(ph[:4] / np.pi, [0.25, 0.75, 1.25, 1.75])
    assert np.allclose(np.cos(th[4:12]), 2 / 3) and np.allclose(ph[4:12] / np.pi, (np.arange(8) + 0.5) / 4)
    assert np.allclose(np.cos(th[12:20]), 1 / 3) and np.allclose(ph[12:20] / np.pi, np.arange(8) / 4)   # belt, unshifted
    assert np.allclose(np.cos(th[20:28]), 0) and np.allclose(ph[20:28] / np.pi, (np.arange(8) + 0.5) / 4)
    for nside in (1, 2, 4, 8):
        rt = H.ring_table(nside)
        assert len(rt) == 4 * nside - 1 and sum(r[0] for r in rt) == 12 * nside * nside
        assert [r[1] for r in rt] == list(np.cumsum([0] + [r[0] for r in rt[:-1]]))
        z = np.array([r[3] for r in rt])
        assert np.allclose(z, -z[::-1])  # north/south mirror symmetry


def test_healpix_oracle_transform_properties():
    from oracle import healpix_ref as H

    nside, L = 8, 12
    rng = np.random.default_rng(0)
    Y = H.ylm_matrix(nside, L)
    alm = np.zeros(H.alm_size(L - 1), complex)
    for el in range(L):
        for m in range(el + 1):
            alm[H.alm_index(el, m, L - 1)] = rng.standard_normal() + (1j * rng.standard_normal() if m else 0)
    # healpy index convention and the real-map symmetry of lm_hp2lm
    assert H.alm_index(0, 0, 11) == 0 and H.alm_index(11, 0, 11) == 11 and H.alm_index(1, 1, 11) == 12
    flm = H.lm_hp2lm(alm, L)
    assert np.array_equal(H.lm2lm_hp(flm, L), alm)
    f = H.synthesis_complex(flm, nside, L, Y)
    assert np.abs(f.imag).max() < 1e-13  # real field
    # monopole: equal-area pixels integrate a constant exactly
    assert np.isclose(H.map2alm(np.full(768, 2.5), L - 1, iter=0, Y=Y)[0], 2.5 * np.sqrt(4 * np.pi), rtol=1e-13)
    # Jacobi refinement converges towards the band-limited input (healpy's iter=3 default)
    mp = H.alm2map(alm, nside, Y)
    errs = [np.abs(H.map2alm(mp, L - 1, iter=it, Y=Y) - alm).max() for it in (0, 1, 3)]
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-4
    # Euclidean adjoint pair
    g = rng.standard_normal(768) + 1j * rng.standard_normal(768)
    x = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    assert np.isclose(np.vdot(g, H.synthesis_complex(x, nside, L, Y)), np.vdot(H.adjoint_complex(g, nside, L, Y), x))


def test_host_lm_hp2lm_matches_oracle():
    from oracle import healpix_ref as H
    from pxmcmc_b200 import utils

    L = 9
    rng = np.random.default_rng(1)
    alm = rng.standard_normal(H.alm_size(L - 1)) + 1j * rng.standard_normal(H.alm_size(L - 1))
    assert np.array_equal(utils.lm_hp2lm(alm, L), H.lm_hp2lm(alm, L))
    assert np.array_equal(utils.lm2lm_hp(utils.lm_hp2lm(alm, L), L), alm)
    with pytest.raises(ValueError):
        utils.lm_hp2lm(alm[:-1], L)
    with pytest.raises(ValueError):
        utils.map2alm(np.zeros(13), 3)


@pytest.mark.parametrize("L,B,J_min", [(12, 2.0, 1), (14, 1.5, 2)])
def test_wavelet_coefficients_are_the_defining_inner_products(L, B, J_min):
    """Second, independently written derivation of the s2let / so3 normalisation (oracle/__init__.py "parity status"):
    the coefficient maps of `analysis_px2wav` are compared with the DEFINITION of the scale-discretised transform,
        W^j(w0) = <f, R_w0 psi^j> = int f(w) conj((R_w0 psi^j)(w)) dOmega ,   S(w0) = <f, R_w0 phi> ,
    evaluated by direct quadrature of closed-form harmonics (scipy sph_harm_y) with the wavelets psi^j_lm / phi_l that
    `wavelet_tiling` returns (the arrays the reference itself consumes in pxmcmc/prior.py:120-149).  This fixes, without
    reference to the transform code: the (2 pi)^(-1/2) of the wavelet coefficients versus none for the scaling
    coefficients, the sqrt((2l+1)/8 pi^2) / sqrt((2l+1)/4 pi) normalisations of the tiling, the sample positions of the
    multiresolution grids and the [scaling, wavelets by increasing j] ordering."""
    from scipy.special import sph_harm_y

    rng = np.random.default_rng(L)
    flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    f = ssht_ref.inverse(flm, L, 0).ravel()
    f_wav, f_scal = s2let_ref.analysis_px2wav(f, B, L, J_min)
    phi_l, psi_lm = s2let_ref.wavelet_tiling(B, L, 1, J_min, 0)
    bls = s2let_ref.bandlimits(B, L, J_min)
    # quadrature grid: the integrand f * conj(R psi) has bandlimit 2L - 1 -> exact MW quadrature at bandlimit 2L
    Lq = 2 * L
    fq = ssht_ref.inverse(np.concatenate([flm, np.zeros(Lq * Lq - L * L)]), Lq, 0).ravel()
    wq = R.mw_map_weights(Lq)
    thq, phq = ssht_ref.sample_positions(Lq)
    TH, PH = np.meshgrid(thq, phq, indexing="ij")
    TH, PH = TH.ravel(), PH.ravel()
    Yq = {(el, m): sph_harm_y(el, m, TH, PH) for el in range(L) for m in range(-el, el + 1)}

    def rotated_kernel(k_l0, th0, ph0):
        """(R_w0 k)(w) for an axisymmetric kernel with harmonic coefficients k_{l0}: addition theorem"""
        out = np.zeros(TH.size, dtype=complex)
        for el in range(L):
            if k_l0[el] == 0:
                continue
            c = k_l0[el] * np.sqrt(4 * np.pi / (2 * el + 1))
            for m in range(-el, el + 1):
                out += c * np.conj(sph_harm_y(el, m, th0, ph0)) * Yq[(el, m)]
        return out

    ls = np.arange(L)
    kernels = [phi_l] + [psi_lm[ls * ls + ls, j] for j in range(psi_lm.shape[1])]
    coef = np.concatenate([f_scal, f_wav])
    off = 0
    for Lj, k_l0 in zip(bls, kernels):
        th, ph = ssht_ref.sample_positions(Lj)
        for (t, p) in [(0, 0), (Lj // 2, Lj // 3), (Lj - 1, 2 * Lj - 2), (1, Lj)]:
            direct = np.sum(wq * fq * np.conj(rotated_kernel(k_l0, th[t], ph[p])))
            got = coef[off + t * (2 * Lj - 1) + p]
            assert abs(got - direct) < 1e-11 * max(1.0, abs(direct)), (Lj, t, p, got, direct)
        off += Lj * (2 * Lj - 1)
    assert off == coef.size


def test_wavelet_power_weights_closed_form():
    """S2_Wavelets_L1_Power_Weights (pxmcmc/prior.py:120-149): the powers P = sum_l |psi_l0|^2 are
    sum_l (2l+1)/(8 pi^2) kappa_j(l)^2 resp. sum_l (2l+1)/(4 pi) kappa_0(l)^2 -- evaluated here from the tiling alone and
    compared with the weight vectors of the restated prior"""
    L, B, J = 20, 2.0, 2
    kappa, kappa0 = s2let_ref.tiling_axisym(B, L, J)
    ls = np.arange(L)
    prior = R.S2WaveletsL1PowerWeights("synthesis", None, None, 1.0, L, B, J, eta=1)
    bls = s2let_ref.bandlimits(B, L, J)
    off = 0
    P0 = np.sum((2 * ls + 1) / (4 * np.pi) * kappa0 ** 2)
    Le = int(np.nonzero(kappa0)[0].max()) + 1
    th, _ = ssht_ref.sample_positions(Le)
    assert np.allclose(prior.map_weights[: Le * (2 * Le - 1)].reshape(Le, -1)[:, 0], 2 * np.pi ** 2 / (P0 * Le * (2 * Le - 1)) * np.sin(th))
    off = bls[0] * (2 * bls[0] - 1)
    for j, Lj in zip(range(J, s2let_ref.j_max(L, B) + 1), bls[1:]):
        Pj = np.sum((2 * ls + 1) / (8 * np.pi ** 2) * kappa[j] ** 2)
        peak = int(np.argmax(kappa[j] * np.sqrt(2 * ls + 1)))
        th, _ = ssht_ref.sample_positions(Lj)
        w = prior.map_weights[off: off + Lj * (2 * Lj - 1)].reshape(Lj, -1)[:, 0]
        assert np.allclose(w, 2 * np.pi ** 2 * peak / (Pj * Lj * (2 * Lj - 1)) * np.sin(th), rtol=1e-12), j
        off += Lj * (2 * Lj - 1)


# ------------------------------------------------------------------ algebra behind the carried forms (product: forward.py)
def test_gram_of_the_mw_inverse_transform_is_diagonal_in_m(rng):
    """What the harmonic (Gram) form of the data-fidelity gradient rests on (pxmcmc_b200/forward.py):
    with a noise level that is constant along every ring, A_inv^dagger diag(w) A_inv does not couple different
    azimuthal orders, and on each order it is (2L-1) Lambda^T diag(w_t) Lambda with the REAL colatitude factor
    Lambda^m[t, l] = A_inv e_lm sampled at phi = 0 -- checked with the oracle's own transforms, no GPU involved"""
    L = 7
    n = 2 * L - 1
    A = np.stack([ssht_ref.inverse(np.eye(L * L, dtype=complex)[i], L).ravel() for i in range(L * L)], axis=1)  # [npix, L^2]
    w_ring = 0.3 + rng.random(L)
    W = np.repeat(w_ring, n)
    G = A.conj().T @ (W[:, None] * A)
    ms = np.array([ssht_ref.ind2elm(i)[1] for i in range(L * L)])
    off = G[ms[:, None] != ms[None, :]]
    assert np.max(np.abs(off)) < 1e-12 * np.max(np.abs(G))
    for m in range(-(L - 1), L):
        idx = np.nonzero(ms == m)[0]
        lam = A.reshape(L, n, L * L)[:, 0, :][:, idx]  # phi = 0 column: e^{i m 0} = 1
        assert np.max(np.abs(lam.imag)) < 1e-13
        assert np.allclose(G[np.ix_(idx, idx)], n * lam.real.T @ (w_ring[:, None] * lam.real), atol=1e-12)
    # and the adjoint used on the way back is the plain conjugate transpose (Euclidean adjoint, as the reference's)
    x, y = rng.standard_normal(L * L) + 0j, rng.standard_normal(L * n) + 0j
    assert np.isclose(np.vdot(y, ssht_ref.inverse(x, L).ravel()), np.vdot(ssht_ref.inverse_adjoint(y.reshape(L, n), L), x))


def test_wavelet_synthesis_and_its_adjoint_are_real_operators(rng):
    """What MYULA(real_pairs=True) rests on: the axisymmetric wavelet synthesis and its adjoint are complex-linear and
    map real fields to real fields, so two real chains can travel as the real and imaginary part of one complex chain"""
    L, B, J = 12, 2.0, 1
    t = R.WaveletTransform(L, B, J)
    xa, xb = rng.standard_normal(t.ncoefs), rng.standard_normal(t.ncoefs)
    pa, pb = t.inverse(xa.astype(complex)), t.inverse(xb.astype(complex))
    assert np.max(np.abs(pa.imag)) < 1e-13 * np.max(np.abs(pa.real))
    assert rel_l2(t.inverse(xa + 1j * xb), pa.real + 1j * pb.real) < 1e-13
    ya, yb = rng.standard_normal(L * (2 * L - 1)), rng.standard_normal(L * (2 * L - 1))
    ca, cb = t.inverse_adjoint(ya.astype(complex)), t.inverse_adjoint(yb.astype(complex))
    assert np.max(np.abs(ca.imag)) < 1e-13 * np.max(np.abs(ca.real))
    assert rel_l2(t.inverse_adjoint(ya + 1j * yb), ca.real + 1j * cb.real) < 1e-13
