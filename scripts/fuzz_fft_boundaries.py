"""Bandlimits at the Bluestein-length boundaries (M switches 256 -> 512 at L = 65, 512 -> 1024 at L = 129) with the
persistent staged ring FFT forced on (mode 3): synthesis pair against the CPU oracle, analysis pair by dot-test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pxmcmc_ref as R
from pxmcmc_b200 import device as D
from pxmcmc_b200._lib import lib

rng = np.random.default_rng(5)
rel = lambda a, b: float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b)))
for L in (64, 65, 127, 128, 129, 130, 255, 256):
    B, J, nb = 2.0, 2, 3
    plan = D.WaveletPlan(L, B, J, nb)
    t = R.WaveletTransform(L, B, J)
    coef = rng.standard_normal((nb, plan.ncoefs)) + 1j * rng.standard_normal((nb, plan.ncoefs))
    pix = rng.standard_normal((nb, plan.npix)) + 1j * rng.standard_normal((nb, plan.npix))
    out = []
    for mode in (1, 3):
        lib.pxm_debug_set_fft_multipass(mode)
        s = D.to_host(plan.synthesis(D.to_dev_c(coef)))
        sa = D.to_host(plan.synthesis_adjoint(D.to_dev_c(pix)))
        a = D.to_host(plan.analysis(D.to_dev_c(pix)))
        aa = D.to_host(plan.analysis_adjoint(D.to_dev_c(coef)))
        out.append((s, sa, a, aa))
    lib.pxm_debug_set_fft_multipass(0)
    e_or = max(rel(out[1][0][1], t.inverse(coef[1])), rel(out[1][1][2], t.inverse_adjoint(pix[2])))
    e_13 = max(rel(x, y) for x, y in zip(out[1], out[0]))
    d = abs(np.vdot(coef[0], out[1][2][0]) - np.vdot(out[1][3][0], pix[0])) / abs(np.vdot(coef[0], out[1][2][0]))
    print(f"L={L:3d} scales={plan.bandlimits}: mode 3 vs oracle {e_or:.2e}, mode 3 vs mode 1 {e_13:.2e}, analysis dot-test {d:.2e}"
          + ("   <-- FAIL" if max(e_or, e_13, d) > 1e-10 else ""), flush=True)
