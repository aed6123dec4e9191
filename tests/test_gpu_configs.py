"""GPU parity at the BASELINE.json configurations (run on the B200 box: `pytest -m gpu`).

One injected-noise iteration of every named configuration, built exactly as SURVEY.md 8(d)
specifies, through the public classes (= the C ABI of libpxmcmc_b200.so), against the CPU
oracle on the same inputs; plus the four SHT primitives at spin 0 / 2 at the BASELINE
bandlimits 256 and 512.  Bars (BASELINE.json north_star): 1e-10 relative L2 in FP64 on
the chain state, the predictions and the gradient; prox support and sign pattern identical.

The oracle is O(L^3) numpy on one core (1 s per SHT at L=256, 9 s at L=512): independent
oracle evaluations of one test run concurrently in worker processes.
"""
import os
from concurrent.futures import ProcessPoolExecutor
import multiprocessing as mp

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-10


# ------------------------------------------------------------------ oracle workers (CPU processes)
def _limit_threads():
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:  # noqa: BLE001
        pass


def _w_sht(args):
    name, x, L, spin = args
    _limit_threads()
    from oracle import ssht_ref

    return getattr(ssht_ref, name)(x, L, spin)


def _w_wav(args):
    name, x, L, B, J = args
    _limit_threads()
    from oracle import pxmcmc_ref as R

    return getattr(R.WaveletTransform(L, B, J), name)(x)


def _w_wl_gradg(args):
    """gradg of config 4 (pxmcmc/forward.py:66-72): Psi^dagger Phi^dagger invcov (preds - data)"""
    preds, data, invcov, L, B, J, mask, ngal = args
    _limit_threads()
    from oracle import pxmcmc_ref as R

    wl = R.WeakLensing(L, mask=mask, ngal=ngal)
    return R.WaveletTransform(L, B, J).inverse_adjoint(wl.adjoint(invcov * (preds - data)))


def _w_wl_forward(args):
    X, L, B, J, mask, ngal = args
    _limit_threads()
    from oracle import pxmcmc_ref as R

    return R.WeakLensing(L, mask=mask, ngal=ngal).forward(R.WaveletTransform(L, B, J).inverse(X))


def _w_skrock(args):
    Acsr, data, sig, L, B, J, delta, lmda, s, X, Z = args
    _limit_threads()
    from oracle import pxmcmc_ref as R

    t = R.WaveletTransform(L, B, J)
    op = R.ForwardOperator(data, sig, "synthesis", t, R.PathIntegral(Acsr), t.ncoefs)
    prior = R.S2WaveletsL1PowerWeights("synthesis", t.inverse, t.inverse_adjoint, lmda, L, B, J, eta=1)
    return R.skrock_step(op, prior, delta, lmda, s, X, Z), prior.T


def _w_myula_synthesis(args):
    data, sig, L, B, J, delta, lmda, mu, X, preds, w = args
    _limit_threads()
    from oracle import pxmcmc_ref as R

    t = R.WaveletTransform(L, B, J)
    op = R.ForwardOperator(data, sig, "synthesis", t, R.IdentityMeasurement(data.size, data.size), t.ncoefs)
    prior = R.S2WaveletsL1("synthesis", t.inverse, t.inverse_adjoint, lmda * mu, L, B, J)
    gradg = op.calc_gradg(preds)
    prox = prior.proxf(X)
    Xn = R.myula_step(X, prox, gradg, delta, lmda, w)
    return gradg, prox, Xn, op.forward(Xn), R.logpi(op, prior, mu, X, preds)


def _w_prox_analysis(args):
    X, T, L, B, J = args
    _limit_threads()
    from oracle import pxmcmc_ref as R

    t = R.WaveletTransform(L, B, J)
    return R.L1("analysis", t.inverse, t.inverse_adjoint, T).proxf(X)


@pytest.fixture(scope="module")
def pool():
    n = max(1, min(8, (os.cpu_count() or 2) - 1))
    with ProcessPoolExecutor(n, mp_context=mp.get_context("spawn")) as ex:
        yield ex


@pytest.fixture(scope="module")
def px():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pxmcmc_b200 import forward, mcmc, measurements, prior, sht, transforms, utils

    class NS:
        pass

    ns = NS()
    ns.forward, ns.mcmc, ns.measurements, ns.prior, ns.sht, ns.transforms, ns.utils = (
        forward, mcmc, measurements, prior, sht, transforms, utils)
    return ns


# ------------------------------------------------------------------ synthetic inputs
def synthetic_flm(L, seed=20240):
    """conjugate-symmetric flm ~ N(0, C_l), C_l = (1+l)^-2.5"""
    rng = np.random.default_rng(seed)
    flm = np.zeros(L * L, dtype=complex)
    for el in range(L):
        amp = (1.0 + el) ** -1.25
        flm[el * el + el] = amp * rng.standard_normal()
        m = np.arange(1, el + 1)
        a = amp * (rng.standard_normal(el) + 1j * rng.standard_normal(el)) / np.sqrt(2)
        flm[el * el + el + m] = a
        flm[el * el + el - m] = (-1.0) ** m * np.conj(a)
    return flm


def unit_rms_map(px, L):
    """data = A_inv(L,0) flm kept complex (as on the reference's HEALPix path), unit RMS.  It is an INPUT of
    both sides, so it may come from the device transform."""
    d = px.sht.inverse(synthetic_flm(L), L).ravel()
    return d / np.sqrt(np.mean(np.abs(d) ** 2))


def same_support_and_signs(a, b):
    return (np.array_equal(a == 0, b == 0) and np.array_equal(np.sign(a.real), np.sign(b.real))
            and np.array_equal(np.sign(a.imag), np.sign(b.imag)))


# ------------------------------------------------------------------ SHT primitives at the BASELINE bandlimits
@pytest.mark.parametrize("L", [256, 512])
def test_sht_primitives_at_baseline_bandlimits(px, pool, L):
    """pyssht.forward / inverse / inverse_adjoint / forward_adjoint, spin 0 and 2 (measurements.py:223-239)"""
    rng = np.random.default_rng(L)
    f = rng.standard_normal((L, 2 * L - 1)) + 1j * rng.standard_normal((L, 2 * L - 1))
    jobs, ours = [], []
    for spin in (0, 2):
        flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
        flm[: spin * spin] = 0
        for name, x in (("inverse", flm), ("forward", f), ("inverse_adjoint", f), ("forward_adjoint", flm)):
            jobs.append((name, x, L, spin))
            ours.append(np.asarray(getattr(px.sht, name)(x, L, Spin=spin)).ravel())
    for job, mine, ref in zip(jobs, ours, pool.map(_w_sht, jobs)):
        err = rel_l2(mine, ref)
        assert err < TOL, f"{job[0]} L={L} spin={job[3]}: rel-L2 {err:.2e}"


@pytest.mark.parametrize("L,B", [(256, 1.5), (512, 2.0)])
def test_wavelet_operators_at_baseline_bandlimits(px, pool, L, B):
    """the four pys2let calls of transforms.py:102-154 at the bandlimits of configs 2/5 and 4"""
    J = 2
    t = px.transforms.SphericalWaveletTransform(L, B, J)
    rng = np.random.default_rng(L + 1)
    pix = rng.standard_normal(L * (2 * L - 1)) + 1j * rng.standard_normal(L * (2 * L - 1))
    coef = rng.standard_normal(t.ncoefs) + 1j * rng.standard_normal(t.ncoefs)
    jobs = [("forward", pix, L, B, J), ("inverse", coef, L, B, J), ("inverse_adjoint", pix, L, B, J),
            ("forward_adjoint", coef, L, B, J)]
    ours = [getattr(t, j[0])(j[1]) for j in jobs]
    for job, mine, ref in zip(jobs, ours, pool.map(_w_wav, jobs)):
        err = rel_l2(mine, ref)
        assert err < TOL, f"{job[0]} L={L} B={B}: rel-L2 {err:.2e}"


# ------------------------------------------------------------------ config 1
def test_config1_myula_L32_iterations(px, pool):
    """earthtopography MYULA, S2_Wavelets_L1, wavelet synthesis, L=32 B=1.5 J_min=2, sig_d=1, lmda=delta=1e-6, mu=1
    (experiments/earthtopography/main.py:50,123-133): five consecutive iterations with injected noise"""
    L, B, J = 32, 1.5, 2
    data = unit_rms_map(px, L)
    op = px.forward.SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, J)
    assert op.nparams == 6493
    prm = px.mcmc.PxMCMCParams(delta=1e-6, lmda=1e-6, mu=1.0, nsamples=1, verbosity=0, track=[])
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, prm.lmda * prm.mu, L=L, B=B, J_min=J)
    m = px.mcmc.MYULA(op, reg, prm)
    rng = np.random.default_rng(1)
    X = rng.laplace(size=op.nparams).astype(complex)
    P = op.forward(X)
    Xo, Po = X.copy(), P.copy()
    Xd, Pd = m._state(X), m._state(P)
    for it in range(5):
        np.random.seed(100 + it)
        w = np.random.randn(op.nparams)
        np.random.seed(100 + it)
        Xd, Pd = m.iterate(Xd, Pd)
        gradg, prox, Xo, Po, _ = _w_myula_synthesis((data, 1.0, L, B, J, 1e-6, 1e-6, 1.0, Xo, Po, w))
        assert rel_l2(Xd.cpu().numpy()[0], Xo) < TOL and rel_l2(Pd.cpu().numpy()[0], Po) < TOL


# ------------------------------------------------------------------ config 5 (and the bench workload)
def test_config5_myula_L256_chain_batch(px, pool):
    """64-chain MYULA sweep of the bench (L=256, B=1.5, synthesis, S2_Wavelets_L1): a 4-chain slice run as ONE batch,
    every chain against its own oracle iteration with the same injected noise (mcmc.py:157-164)"""
    L, B, J, nch = 256, 1.5, 2, 4
    data = unit_rms_map(px, L)
    op = px.forward.SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, J, nchains=nch)
    assert op.nparams == 398342
    prm = px.mcmc.PxMCMCParams(delta=1e-6, lmda=1e-6, mu=1.0, nsamples=1, verbosity=0, track=[])
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, prm.lmda * prm.mu, L=L, B=B, J_min=J)
    m = px.mcmc.MYULA(op, reg, prm, nchains=nch)
    rng = np.random.default_rng(2)
    X = rng.laplace(size=(nch, op.nparams)) * np.array([1.0, 1e-3, 1e-6, 1e-5])[:, None]  # thresholds bite on chains 2, 3
    X = X.astype(complex)
    Xd = m._state(X)
    Pd = op.forward(Xd)
    P = Pd.cpu().numpy()
    np.random.seed(77)
    w = np.random.randn(nch * op.nparams).reshape(nch, op.nparams)
    np.random.seed(77)
    gd = op.calc_gradg(Pd).cpu().numpy()
    proxd = reg.proxf(Xd).cpu().numpy()
    lp, l2, pr = m._logpi_dev(Xd, Pd)
    Xn, Pn = m.iterate(Xd, Pd)
    Xn, Pn = Xn.cpu().numpy(), Pn.cpu().numpy()
    jobs = [(data, 1.0, L, B, J, 1e-6, 1e-6, 1.0, X[c], P[c], w[c]) for c in range(nch)]
    for c, (gradg, prox, Xo, Po, (lpo, l2o, pro)) in enumerate(pool.map(_w_myula_synthesis, jobs)):
        assert rel_l2(gd[c], gradg) < TOL, f"gradg chain {c}"
        assert same_support_and_signs(proxd[c], prox), f"prox support / signs chain {c}"
        assert rel_l2(proxd[c], prox) < 1e-14
        assert rel_l2(Xn[c], Xo) < TOL, f"state chain {c}: {rel_l2(Xn[c], Xo):.2e}"
        assert rel_l2(Pn[c], Po) < TOL, f"predictions chain {c}: {rel_l2(Pn[c], Po):.2e}"
        assert np.isclose(lp[c], lpo, rtol=1e-10) and np.isclose(l2[c], l2o, rtol=1e-10) and np.isclose(pr[c], pro, rtol=1e-12)
    assert 0 < np.count_nonzero(proxd[2] == 0) < op.nparams  # the support test is not vacuous
    # the form MYULA.run and bench.py carry the predictions in (ring coefficients of the image: the pixel-side ring-FFT pair
    # of consecutive iterations cancels) against the same oracle iterations
    from pxmcmc_b200.forward import RingPreds

    Pr = m._initial_preds(Xd)
    assert isinstance(Pr, RingPreds)
    np.random.seed(77)
    Xr, Pr = m.iterate(Xd, Pr)
    Xr, Prp = Xr.cpu().numpy(), Pr.pixels().cpu().numpy()
    for c in range(nch):
        assert rel_l2(Xr[c], Xn[c]) < 1e-12 and rel_l2(Prp[c], Pn[c]) < 1e-12, f"ring-carried predictions, chain {c}"


def test_config5_real_chain_pairs_L256(px, pool):
    """the same sweep on REAL data (the earth-topography drivers' case) with MYULA(real_pairs=True): two real chains
    travel as one complex chain on the device; every real chain against its own oracle iteration (which, like the
    reference, carries it as a complex array with a zero imaginary part) with the same injected noise"""
    L, B, J, nch = 256, 1.5, 2, 4
    data = np.ascontiguousarray(unit_rms_map(px, L).real)
    data /= np.sqrt(np.mean(data ** 2))
    op = px.forward.SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, J, nchains=nch)
    prm = px.mcmc.PxMCMCParams(delta=1e-6, lmda=1e-6, mu=1.0, nsamples=1, verbosity=0, track=[])
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, prm.lmda * prm.mu, L=L, B=B, J_min=J)
    m = px.mcmc.MYULA(op, reg, prm, nchains=nch, real_pairs=True)
    eng = m.engine
    assert eng is not m and eng.nchains == nch // 2
    rng = np.random.default_rng(2)
    X = rng.laplace(size=(nch, op.nparams)) * np.array([1.0, 1e-3, 1e-6, 1e-5])[:, None]
    Xd = m._state(X.astype(complex))
    P = op.forward(Xd).cpu().numpy()
    np.random.seed(77)
    w = np.random.randn(nch * op.nparams).reshape(nch, op.nparams)
    Xp = m.pack(Xd)
    Pp = eng._initial_preds(Xp)
    np.random.seed(77)
    Xn, Pn = eng.iterate(Xp, Pp)
    Xn, Pn = m.unpack(Xn).cpu().numpy(), m.unpack(eng._pix(Pn)).cpu().numpy()
    assert np.all(Xn.imag == 0) and np.all(Pn.imag == 0)
    jobs = [(data, 1.0, L, B, J, 1e-6, 1e-6, 1.0, X[c].astype(complex), P[c], w[c]) for c in range(nch)]
    for c, (gradg, prox, Xo, Po, _) in enumerate(pool.map(_w_myula_synthesis, jobs)):
        assert rel_l2(Xn[c], Xo) < TOL, f"state chain {c}: {rel_l2(Xn[c], Xo):.2e}"
        assert rel_l2(Pn[c], Po) < TOL, f"predictions chain {c}: {rel_l2(Pn[c], Po):.2e}"


# ------------------------------------------------------------------ config 2
def test_config2_pxmala_analysis_L256_iteration(px, pool):
    """earthtopography PxMALA, wavelet analysis prior L1("analysis", Psi, Psi^dagger, T), L=256 B=1.5, complex
    HEALPix-path data: the quantities of one iteration of mcmc.py:230-259 -- proposal with injected noise, both
    proximal maps, both transition terms, log pi, log alpha, accept decision and tuned step size"""
    from oracle import pxmcmc_ref as R

    L, B, J = 256, 1.5, 2
    npix = L * (2 * L - 1)
    data = unit_rms_map(px, L)
    sig, lmda, delta, mu = 0.1, 1e-6, 1e-7, 1.0
    op = px.forward.SphericalWaveletTransformOperator(data, sig, "analysis", L, B, J)
    assert op.nparams == npix
    prm = px.mcmc.PxMCMCParams(delta=delta, lmda=lmda, mu=mu, nsamples=1, verbosity=0, track=[])
    T = lmda * mu * 2e5  # the threshold bites on part of the wavelet coefficients of X
    reg = px.prior.L1("analysis", op.transform.inverse, op.transform.inverse_adjoint, T)
    m = px.mcmc.PxMALA(op, reg, prm, tune_delta=True)
    rng = np.random.default_rng(3)
    Xc = (data.real + 0.05 * rng.standard_normal(npix)).astype(complex)
    np.random.seed(5)
    w = np.random.randn(npix)
    u = np.random.rand()
    # --- device
    Xd = m._state(Xc)
    Pc = op.forward(Xd)
    gc = op.calc_gradg(Pc)
    pc = m._proxf_dev(Xd)
    np.random.seed(5)
    Xp = m._propose_dev(Xd, pc, gc)
    Pp = op.forward(Xp)
    gp = op.calc_gradg(Pp)
    pp = m._proxf_dev(Xp)
    t_cp = m._logtrans_dev(Xd, Xp, pc, gc)
    t_pc = m._logtrans_dev(Xp, Xd, pp, gp)
    lpc = m._logpi_dev(Xd, Pc)[0][0]
    lpp = m._logpi_dev(Xp, Pp)[0][0]
    la = t_pc + lpp - t_cp - lpc
    host = lambda t: t.cpu().numpy()[0]  # noqa: E731
    # --- oracle (the two proximal maps are the O(L^3) part: concurrently; the proposal needs prox(Xc) first)
    ident = R.IdentityMeasurement(npix, npix)
    oop = R.ForwardOperator(data, sig, "analysis", None, ident, npix)
    o_pc = _w_prox_analysis((Xc, T, L, B, J))
    o_gc = oop.calc_gradg(oop.forward(Xc))
    o_Xp = R.myula_step(Xc, o_pc, o_gc, delta, lmda, w)
    o_pp = list(pool.map(_w_prox_analysis, [(o_Xp, T, L, B, J)]))[0]
    o_gp = oop.calc_gradg(oop.forward(o_Xp))
    assert rel_l2(host(pc), o_pc) < TOL and rel_l2(host(gc), o_gc) < TOL
    assert rel_l2(host(Xp), o_Xp) < TOL, f"proposal: {rel_l2(host(Xp), o_Xp):.2e}"
    assert rel_l2(host(pp), o_pp) < TOL and rel_l2(host(gp), o_gp) < TOL
    assert rel_l2(host(pc) - Xc, o_pc - Xc) < 1e-8  # the prox actually moved X: compare the move itself too
    prior0 = R.L1("analysis", None, None, T)
    o_lpc = R.logpi(oop, prior0, mu, Xc, oop.forward(Xc))[0]
    o_lpp = R.logpi(oop, prior0, mu, o_Xp, oop.forward(o_Xp))[0]
    o_tcp = R.pxmala_logtransition(Xc, o_Xp, o_pc, o_gc, delta, lmda)
    o_tpc = R.pxmala_logtransition(o_Xp, Xc, o_pp, o_gp, delta, lmda)
    o_la = o_tpc + o_lpp - o_tcp - o_lpc
    for mine, ref in ((lpc, o_lpc), (lpp, o_lpp), (t_cp, o_tcp), (t_pc, o_tpc)):
        assert abs(mine - ref) <= 1e-9 * abs(ref), (mine, ref)
    assert abs(la - o_la) <= 1e-9 * (abs(o_tpc) + abs(o_lpp) + abs(o_tcp) + abs(o_lpc))
    from pxmcmc_b200.mcmc import _cplx_lt

    acc = _cplx_lt(np.log(u), la)
    o_acc = (np.log(u) < o_la)
    assert bool(acc) == bool(o_acc)
    m.acceptance_trace = [int(acc)]
    m._tune_delta(0)
    assert np.isclose(m.delta, R.pxmala_tune_delta(delta, lmda, int(o_acc), 0), rtol=1e-15)


# ------------------------------------------------------------------ config 3
def test_config3_skrock_pathintegral_L128_step(px, pool):
    """phasevel SKROCK s=10 with PathIntegralOperator: 10^4 x 32 640 great-circle CSR (random end points, seed 7, 160
    points per radian, nearest MW pixel, rows sum to 1), sigma = 0.05, S2_Wavelets_L1_Power_Weights(eta=1), lmda = delta/2
    (experiments/phasevel/main.py:45,140-163): one chain_step from (X, Z) against the memoised oracle recursion"""
    from oracle import greatcircle_ref as G

    L, B, J, s, npaths = 128, 2, 2, 10, 10000
    starts, stops = G.random_endpoints(npaths, seed=7)
    A = G.path_matrix(starts, stops, L)
    assert A.shape == (npaths, 32640) and np.allclose(np.asarray(A.sum(axis=1)).ravel(), 1.0)
    rng = np.random.default_rng(8)
    truth = unit_rms_map(px, L).real
    data = A @ truth + 0.05 * rng.standard_normal(npaths)
    sig = np.full(npaths, 0.05)
    delta = 1e-6
    lmda = delta / 2
    op = px.forward.PathIntegralOperator(A, data, sig, "synthesis", L, B, J)
    assert op.nparams == 76068
    prm = px.mcmc.PxMCMCParams(delta=delta, lmda=lmda, mu=1.0, s=s, nsamples=1, verbosity=0, track=[])
    reg = px.prior.S2_Wavelets_L1_Power_Weights("synthesis", op.transform.inverse, op.transform.inverse_adjoint,
                                               prm.lmda * prm.mu, L=L, B=B, J_min=J, eta=1)
    m = px.mcmc.SKROCK(op, reg, prm)
    X = rng.laplace(size=op.nparams) * 1e-2
    np.random.seed(9)
    Z = np.random.randn(op.nparams)
    np.random.seed(9)
    Xn = m.chain_step(X)
    # forward / adjoint of the path operator alone (warp-per-row CSR SpMV, measurements.py:69-83)
    pixv = rng.standard_normal(32640) + 1j * rng.standard_normal(32640)
    yv = rng.standard_normal(npaths) + 1j * rng.standard_normal(npaths)
    assert rel_l2(op.measurement.forward(pixv), A @ pixv) < 1e-13
    assert rel_l2(op.measurement.adjoint(yv), A.conj().T @ yv) < 1e-13
    Xo, To = list(pool.map(_w_skrock, [(A, data, sig, L, B, J, delta, lmda, s, X, Z)]))[0]
    assert rel_l2(reg.T, To) < 1e-13
    assert np.all(np.isfinite(Xn))
    assert rel_l2(Xn, Xo) < TOL, f"SKROCK step: {rel_l2(Xn, Xo):.2e}"


# ------------------------------------------------------------------ config 4
def wl_mask(L):
    """the bench's mask: equatorial band |90deg - theta| < 10deg and the same band in the frame tilted by the inclination of
    the galactic plane (stand-in for utils.build_mask(L, 10))"""
    th = (2 * np.arange(L) + 1) * np.pi / (2 * L - 1)
    ph = 2 * np.pi * np.arange(2 * L - 1) / (2 * L - 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    y, z = np.sin(T) * np.sin(P), np.cos(T)
    a = np.radians(62.87)
    z2 = -np.sin(a) * y + np.cos(a) * z
    mask = np.ones((L, 2 * L - 1), dtype=bool)
    mask[np.abs(np.degrees(np.arcsin(np.clip(z, -1, 1)))) < 10] = False
    mask[np.abs(np.degrees(np.arcsin(np.clip(z2, -1, 1)))) < 10] = False
    return mask


def test_config4_weaklensing_myula_L512_iteration(px, pool):
    """weaklensing MYULA, spin-2 Kaiser-Squires operator behind a wavelet synthesis at L=512 B=2, masked, ngal=30,
    sig_d = 1/inv_cov, lmda = delta/2 = 5e-7 (experiments/weaklensing/main.py:86-119): one iteration with injected noise,
    through the product path (Phi o Psi composed in harmonic space) AND through the reference's literal composition
    (fuse_harmonic = False); the oracle is the literal composition."""
    from oracle import pxmcmc_ref as R
    from oracle import s2let_ref

    L, B, J = 512, 2.0, 2
    mask = wl_mask(L)
    ngal = np.full((L, 2 * L - 1), 30.0)
    wl = px.measurements.WeakLensing(L, mask=mask, ngal=ngal)
    tr = px.transforms.SphericalWaveletTransform(L, B, J)
    assert tr.ncoefs == 1221796
    rng = np.random.default_rng(11)
    x_true = rng.laplace(size=tr.ncoefs) * 1e-3
    clean = wl.forward(tr.inverse(x_true))
    data = clean + (rng.standard_normal(clean.size) + 1j * rng.standard_normal(clean.size))
    sig = 1.0 / wl.inv_cov
    delta, lmda = 1e-6, 5e-7
    X = (x_true + 1e-4 * rng.laplace(size=tr.ncoefs)).astype(complex)
    np.random.seed(12)
    w = np.random.randn(tr.ncoefs)
    res = {}
    for fuse in (True, False):
        op = px.forward.ForwardOperator(data, sig, "synthesis", transform=tr, measurement=wl, nparams=tr.ncoefs)
        op.fuse_harmonic = fuse
        assert op._fused() == fuse
        prm = px.mcmc.PxMCMCParams(delta=delta, lmda=lmda, mu=1.0, nsamples=1, verbosity=0, track=[])
        reg = px.prior.S2_Wavelets_L1("synthesis", tr.inverse, tr.inverse_adjoint, prm.lmda * prm.mu, L=L, B=B, J_min=J)
        m = px.mcmc.MYULA(op, reg, prm)
        Xd = m._state(X)
        Pd = op.forward(Xd)
        assert Pd.shape == (1, wl.ndata)  # a one-chain batch stays a batch through both compositions
        gd = op.calc_gradg(Pd).cpu().numpy().ravel()
        proxd = reg.proxf(Xd).cpu().numpy().ravel()
        np.random.seed(12)
        Xn, Pn = m.iterate(Xd, Pd)
        res[fuse] = (Pd.cpu().numpy().ravel(), gd, proxd, Xn.cpu().numpy().ravel(), Pn.cpu().numpy().ravel(), op._diag)
    P, gd, proxd, Xn, Pn, diag = res[True]
    # oracle: the three O(L^3) evaluations are independent given the device's X, P and X' (all of them compared below)
    invcov = R.inverse_covariance(data, sig)
    assert rel_l2(diag, invcov) < 1e-15
    f_g = pool.submit(_w_wl_gradg, (P, data, invcov, L, B, J, mask, ngal))
    f_p = pool.submit(_w_wl_forward, (X, L, B, J, mask, ngal))
    f_n = pool.submit(_w_wl_forward, (Xn, L, B, J, mask, ngal))
    To = lmda * np.concatenate([R.mw_map_weights(b) for b in s2let_ref.bandlimits(B, L, J)])
    o_prox = R.soft(X, To)
    o_g, o_P, o_Pn = f_g.result(), f_p.result(), f_n.result()
    o_Xn = R.myula_step(X, o_prox, o_g, delta, lmda, w)
    for fuse in (True, False):
        P, gd, proxd, Xn, Pn, _ = res[fuse]
        tag = "harmonic-space composition" if fuse else "literal composition"
        assert same_support_and_signs(proxd, o_prox) and rel_l2(proxd, o_prox) < 1e-14, tag
        assert rel_l2(gd, o_g) < TOL, f"gradg ({tag}): {rel_l2(gd, o_g):.2e}"
        assert rel_l2(Xn, o_Xn) < TOL, f"state ({tag}): {rel_l2(Xn, o_Xn):.2e}"
        assert rel_l2(Xn - X, o_Xn - X) < 1e-8, tag  # the step itself, not only X + step
        assert rel_l2(P, o_P) < TOL, f"predictions of X ({tag})"
        assert rel_l2(Pn, o_Pn) < TOL, f"predictions of X' ({tag})"
