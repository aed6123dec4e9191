"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/pxmcmc/*.py) on top of the oracle shims.  Runs only in the
build container.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_golden

What the fixtures pin
  * everything the reference's own Python layer computes (soft, flatten order,
    inverse covariance incl. the complex-data rule, gradg, prox, chain_step,
    logpi, PxMALA transition/accept/tuning, SKROCK coefficients and recursion,
    weight vectors of the S2 priors, weak-lensing kernel/mask/cov plumbing) --
    this is real reference output;
  * the composition of those with the ssht/s2let restatement (layout,
    multiresolution bookkeeping).  The absolute values of the transforms
    themselves come from oracle/ssht_ref.py + s2let_ref.py (parity unpinned
    against the absent wheels, see oracle/__init__.py).
"""
import contextlib
import io
import os
import sys

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def _real_field_lm(L, rng):
    flm = np.zeros(L * L, dtype=complex)
    for el in range(L):
        flm[el * el + el] = rng.standard_normal()
        for m in range(1, el + 1):
            a = rng.standard_normal() + 1j * rng.standard_normal()
            flm[el * el + el + m] = a
            flm[el * el + el - m] = (-1) ** m * np.conj(a)
    return flm / (1.0 + np.repeat(np.arange(L), 2 * np.arange(L) + 1)) ** 1.25


def main():
    sys.path.insert(0, ROOT)
    from oracle import refloader

    refloader.load()
    import pys2let
    import pyssht
    from pxmcmc.forward import ForwardOperator, PathIntegralOperator, SphericalWaveletTransformOperator
    from pxmcmc.mcmc import MYULA, SKROCK, PxMALA, PxMCMCParams
    from pxmcmc.measurements import WeakLensing
    from pxmcmc.prior import L1, S2_Wavelets_L1, S2_Wavelets_L1_Power_Weights
    from pxmcmc.transforms import SphericalWaveletTransform
    from pxmcmc.utils import flatten_mlm, mw_map_weights, soft

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20240)
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- 1. soft thresholding (utils.py:55-67) -------------------------------
    xr = rng.standard_normal(257) * 2
    xc = rng.standard_normal(257) + 1j * rng.standard_normal(257)
    xc[:3] = [0, 1 + 1j, 0.5 - 0.5j]
    tv = np.abs(rng.standard_normal(257)) * 0.7
    xr[5], tv[5] = 0.25, 0.25  # |x| == T is zeroed (inclusive)
    np.savez(
        os.path.join(OUT, "ref_soft.npz"),
        xr=xr, xc=xc, tv=tv,
        soft_r_scalar=soft(xr, 0.8), soft_r_vec=soft(xr, tv),
        soft_c_scalar=soft(xc, 1.0), soft_c_vec=soft(xc, tv),
    )

    # ---- 2. wavelet transform wrappers (transforms.py) -----------------------
    for tag, (L, B, J_min) in {"L10B2": (10, 2, 2), "L16B1p5": (16, 1.5, 2)}.items():
        tr = SphericalWaveletTransform(L, B, J_min)
        x_pix = pys2let.alm2map_mw(_real_field_lm(L, rng), L, 0)
        x_pix_c = x_pix + 1j * rng.standard_normal(x_pix.size) * 0.1
        x_coef = rng.standard_normal(tr.ncoefs) + 1j * rng.standard_normal(tr.ncoefs)
        np.savez(
            os.path.join(OUT, f"ref_wavelet_{tag}.npz"),
            L=L, B=B, J_min=J_min, nscal=tr.nscal, nwav=tr.nwav,
            x_pix=x_pix_c, x_coef=x_coef,
            forward=tr.forward(x_pix_c), inverse=tr.inverse(x_coef),
            inverse_adjoint=tr.inverse_adjoint(x_pix_c), forward_adjoint=tr.forward_adjoint(x_coef),
            s2_T=S2_Wavelets_L1("synthesis", None, None, 1.0, L, B, J_min).T,
            s2pw_T=S2_Wavelets_L1_Power_Weights("synthesis", None, None, 1.0, L, B, J_min, eta=1).T,
            s2pw_w=S2_Wavelets_L1_Power_Weights("synthesis", None, None, 1.0, L, B, J_min, eta=1).map_weights,
            mw_weights=mw_map_weights(L),
        )

    # ---- 3. pyssht-level primitives used by WeakLensing ----------------------
    L = 12
    flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    f = rng.standard_normal((L, 2 * L - 1)) + 1j * rng.standard_normal((L, 2 * L - 1))
    d = {"L": L, "flm": flm, "f": f}
    for s in (0, 2):
        fl = flm.copy()
        fl[: s * s] = 0
        d[f"flm_s{s}"] = fl
        d[f"inverse_s{s}"] = pyssht.inverse(fl, L, Spin=s)
        d[f"forward_s{s}"] = pyssht.forward(f, L, Spin=s)
        d[f"inverse_adjoint_s{s}"] = pyssht.inverse_adjoint(f, L, Spin=s)
        d[f"forward_adjoint_s{s}"] = pyssht.forward_adjoint(fl, L, Spin=s)
    np.savez(os.path.join(OUT, "ref_sht_L12.npz"), **d)

    # ---- 4. MYULA, synthesis, S2_Wavelets_L1 (config 1 shape at L=10) --------
    def run_myula(tag, L, B, J_min, complex_data, sig_vec, seed):
        data = pys2let.alm2map_mw(_real_field_lm(L, rng), L, 0)
        data = data / np.sqrt(np.mean(np.abs(data) ** 2))
        if not complex_data:
            data = data.real.copy()
        sig_d = np.full(data.size, 0.1) + 0.05 * rng.random(data.size) if sig_vec else 0.1
        op = SphericalWaveletTransformOperator(data, sig_d, "synthesis", L, B, J_min)
        p = PxMCMCParams(nsamples=4, nburn=2, ngap=3, delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0,
                         track=["logposterior", "L2", "prior", "chain", "predictions"])
        reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint,
                             p.lmda * p.mu, L=L, B=B, J_min=J_min)
        np.random.seed(seed)
        m = MYULA(op, reg, p)
        with quiet, np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m.run()
        np.savez(os.path.join(OUT, f"ref_myula_{tag}.npz"), L=L, B=B, J_min=J_min, data=data,
                 sig_d=sig_d, seed=seed, nsamples=4, nburn=2, ngap=3, delta=1e-6, lmda=1e-6, mu=1.0,
                 chain=m.chain, logPi=m.logPi, L2s=m.L2s, priors=m.priors, preds=m.preds)

    run_myula("L10_complex", 10, 2, 2, True, False, 11)
    run_myula("L10_real_sigvec", 10, 2, 2, False, True, 12)
    run_myula("L16B1p5_complex", 16, 1.5, 2, True, True, 13)

    # ---- 5. one explicit MYULA iteration with stored noise -------------------
    L, B, J_min = 10, 2, 2
    data = pys2let.alm2map_mw(_real_field_lm(L, rng), L, 0)
    op = SphericalWaveletTransformOperator(data, 0.2, "synthesis", L, B, J_min)
    p = PxMCMCParams(delta=2e-6, lmda=1e-6, mu=2.0, verbosity=0, nsamples=1)
    reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, p.lmda * p.mu, L=L, B=B, J_min=J_min)
    m = MYULA(op, reg, p)
    X = rng.laplace(size=op.nparams) + 1j * rng.laplace(size=op.nparams) * 0.1
    preds = op.forward(X)
    gradg = op.calc_gradg(preds)
    prox = reg.proxf(X)
    np.random.seed(5)
    w = np.random.randn(op.nparams)
    np.random.seed(5)
    Xn = m.chain_step(X, prox, gradg)
    lp, l2, pr = m.logpi(X, preds)
    np.savez(os.path.join(OUT, "ref_myula_step_L10.npz"), L=L, B=B, J_min=J_min, data=data, sig_d=0.2,
             delta=2e-6, lmda=1e-6, mu=2.0, X=X, preds=preds, gradg=gradg, prox=prox, w=w, Xn=Xn,
             invcov=op.invcov.diagonal(), logpi=lp, L2=l2, prior=pr, T=reg.T)

    # ---- 6. PxMALA, analysis setting, L1 prior (config 2 shape at L=10) ------
    data = pys2let.alm2map_mw(_real_field_lm(L, rng), L, 0)
    data = data / np.sqrt(np.mean(np.abs(data) ** 2))
    op = SphericalWaveletTransformOperator(data, 0.1, "analysis", L, B, J_min)
    p = PxMCMCParams(nsamples=6, nburn=2, ngap=2, delta=1e-7, lmda=1e-6, mu=1.0, verbosity=0,
                     track=["logposterior", "L2", "prior", "chain", "predictions"])
    reg = L1("analysis", op.transform.inverse, op.transform.inverse_adjoint, p.lmda * p.mu)
    np.random.seed(21)
    m = PxMALA(op, reg, p, tune_delta=True)
    import warnings
    with quiet, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m.run()
    # one explicit transition evaluation
    X1 = rng.standard_normal(op.nparams) + 0j
    X2 = X1 + 1e-3 * rng.standard_normal(op.nparams)
    pf = reg.proxf(X1)
    gg = op.calc_gradg(op.forward(X1))
    np.savez(os.path.join(OUT, "ref_pxmala_L10.npz"), L=L, B=B, J_min=J_min, data=data, sig_d=0.1, seed=21,
             nsamples=6, nburn=2, ngap=2, delta=1e-7, lmda=1e-6, mu=1.0,
             chain=m.chain, logPi=m.logPi, L2s=m.L2s, priors=m.priors, preds=m.preds,
             acceptance_trace=np.array(m.acceptance_trace), deltas_trace=np.array(m.deltas_trace),
             X1=X1, X2=X2, proxf=pf, gradg=gg, logtrans=m.calc_logtransition(X1, X2, pf, gg))

    # ---- 7. SKROCK (coefficients + one chain_step; reference recursion) ------
    Ls = 10  # (L=8,B=2 makes all wavelet maps the same shape, which crashes prior.py:149)
    data = pys2let.alm2map_mw(_real_field_lm(Ls, rng), Ls, 0).real.copy()
    A = sparse.random(40, Ls * (2 * Ls - 1), density=0.08, random_state=3, format="csr")
    A = sparse.diags(1.0 / np.maximum(np.asarray(A.sum(axis=1)).ravel(), 1e-3)) @ A
    ydata = A @ data + 0.01 * rng.standard_normal(40)
    sig = np.full(40, 0.05)
    op = PathIntegralOperator(A.tocsr(), ydata, sig, "synthesis", Ls, 2, 2)
    p = PxMCMCParams(delta=1e-6, lmda=5e-7, mu=1.0, s=3, verbosity=0, nsamples=1)
    reg = S2_Wavelets_L1_Power_Weights("synthesis", op.transform.inverse, op.transform.inverse_adjoint,
                                        p.lmda * p.mu, L=Ls, B=2, J_min=2, eta=1)
    m = SKROCK(op, reg, p)
    X = rng.laplace(size=op.nparams) * 0.1
    np.random.seed(31)
    Z = np.random.randn(op.nparams)
    np.random.seed(31)
    Xn = m.chain_step(X)
    m5 = SKROCK(op, reg, PxMCMCParams(s=5, verbosity=0, nsamples=1))
    np.savez(os.path.join(OUT, "ref_skrock_L10.npz"), L=Ls, B=2, J_min=2, data=ydata, sig_d=sig,
             A_data=A.tocsr().data, A_indices=A.tocsr().indices, A_indptr=A.tocsr().indptr, A_shape=np.array(A.shape),
             delta=1e-6, lmda=5e-7, mu=1.0, s=3, X=X, Z=Z, Xn=Xn, mus=m.mus, nus=m.nus, ks=m.ks,
             omega_0=m.omega_0, omega_1=m.omega_1, mus5=m5.mus, nus5=m5.nus, ks5=m5.ks,
             omega5_0=m5.omega_0, omega5_1=m5.omega_1,
             preds=op.forward(X), gradg=op.calc_gradg(op.forward(X)), T=reg.T,
             prior=reg.prior(X), gradlogpi=m._gradlogpi(X))

    # ---- 8. weak lensing (measurements.py:185-304) ---------------------------
    Lw = 12
    mask = (rng.random((Lw, 2 * Lw - 1)) > 0.4).astype(int)
    ngal = rng.integers(5, 40, size=(Lw, 2 * Lw - 1)).astype(float)
    wl = WeakLensing(Lw, mask=mask, ngal=ngal)
    kappa = rng.standard_normal(Lw * (2 * Lw - 1)) + 1j * rng.standard_normal(Lw * (2 * Lw - 1)) * 0.2
    gam = rng.standard_normal(int(mask.sum())) + 1j * rng.standard_normal(int(mask.sum()))
    wl0 = WeakLensing(Lw)
    # full forward operator as in experiments/weaklensing/main.py:92-107
    tr = SphericalWaveletTransform(Lw, 2, 2)
    gdata = wl.forward(kappa) + 0.1 * (rng.standard_normal(wl.inv_cov.size) + 1j * rng.standard_normal(wl.inv_cov.size))
    fo = ForwardOperator(gdata, 1 / wl.inv_cov, "synthesis", transform=tr, measurement=wl, nparams=tr.ncoefs)
    Xw = rng.laplace(size=tr.ncoefs) * 0.1 + 0j
    pw = fo.forward(Xw)
    np.savez(os.path.join(OUT, "ref_weaklensing_L12.npz"), L=Lw, mask=mask, ngal=ngal, kappa=kappa, gamma=gam,
             kernel=wl.harmonic_kernel, inv_cov=wl.inv_cov, forward=wl.forward(kappa), adjoint=wl.adjoint(gam),
             forward_nomask=wl0.forward(kappa), adjoint_nomask=wl0.adjoint(kappa),
             gdata=gdata, X=Xw, op_forward=pw, op_gradg=fo.calc_gradg(pw), op_invcov=fo.invcov.diagonal())

    # ---- 9. flatten layout known answer (tests/test_utils.py:8-16) -----------
    w9 = np.ones((861, 9))
    for i in range(9):
        w9[:, i] += i
    np.savez(os.path.join(OUT, "ref_flatten.npz"), flat=flatten_mlm(w9, np.zeros(861)))
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn}: {os.path.getsize(os.path.join(OUT, fn))} B")


if __name__ == "__main__":
    main()
