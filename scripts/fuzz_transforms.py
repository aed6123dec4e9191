"""Randomised sweep of (L, B, J_min, nchains): all four wavelet operators and the spin-0/2 SHTs against the CPU oracle.
A development check (the fixed cases live in tests/); prints the worst relative error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pxmcmc_ref as R, ssht_ref
from pxmcmc_b200 import device as D

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
rel = lambda a, b: float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))
cases = [(2, 2.0, 0), (3, 2.0, 1), (4, 3.0, 0), (5, 1.5, 2), (33, 2.0, 2), (40, 1.3, 3), (17, 4.0, 1)]
for _ in range(14):
    L = int(rng.integers(2, 48))
    B = float(rng.choice([1.3, 1.5, 2.0, 2.5, 3.0]))
    jm = int(np.ceil(np.log(L) / np.log(B)))
    cases.append((L, B, int(rng.integers(0, max(jm, 1)))))
for L, B, J in cases:
    nb = int(rng.integers(1, 4))
    try:
        t = R.WaveletTransform(L, B, J)
    except Exception as e:  # noqa: BLE001
        print(f"L={L} B={B} J_min={J}: oracle rejects ({e})"); continue
    plan = D.WaveletPlan(L, B, J, nb)
    coef = rng.standard_normal((nb, plan.ncoefs)) + 1j * rng.standard_normal((nb, plan.ncoefs))
    pix = rng.standard_normal((nb, plan.npix)) + 1j * rng.standard_normal((nb, plan.npix))
    errs = []
    got = {"synthesis": D.to_host(plan.synthesis(D.to_dev_c(coef))), "synthesis_adjoint": D.to_host(plan.synthesis_adjoint(D.to_dev_c(pix))),
           "analysis": D.to_host(plan.analysis(D.to_dev_c(pix))), "analysis_adjoint": D.to_host(plan.analysis_adjoint(D.to_dev_c(coef)))}
    for c in range(nb):
        errs += [rel(got["synthesis"][c], t.inverse(coef[c])), rel(got["synthesis_adjoint"][c], t.inverse_adjoint(pix[c])),
                 rel(got["analysis"][c], t.forward(pix[c])), rel(got["analysis_adjoint"][c], t.forward_adjoint(coef[c]))]
    for spin in (0, 2):
        if L <= spin:
            continue
        sp = D.ShtPlan(L, spin, 1)
        flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
        flm[: spin * spin] = 0
        errs += [rel(D.to_host(sp.inverse(D.to_dev_c(flm))), ssht_ref.inverse(flm, L, spin)),
                 rel(D.to_host(sp.forward(D.to_dev_c(pix[0]))), ssht_ref.forward(pix[0].reshape(L, 2 * L - 1), L, spin)),
                 rel(D.to_host(sp.inverse_adjoint(D.to_dev_c(pix[0]))), ssht_ref.inverse_adjoint(pix[0].reshape(L, 2 * L - 1), L, spin)),
                 rel(D.to_host(sp.forward_adjoint(D.to_dev_c(flm))), ssht_ref.forward_adjoint(flm, L, spin))]
    e = max(errs)
    worst = max(worst, e)
    print(f"L={L:3d} B={B} J_min={J} chains={nb} scales={plan.bandlimits}: max rel {e:.2e}" + ("   <-- FAIL" if e > 1e-10 else ""), flush=True)
print("worst", worst)
