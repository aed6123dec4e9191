"""Randomised sweep of (L, B, J_min, nchains): MYULA iterations with the predictions carried as pixels (the reference's
literal composition), as ring coefficients, as harmonic coefficients (Gram form) and -- real data -- as packed chain
pairs must agree.  Small and odd shapes are where tile padding and row-group logic would break.  A development check
(the fixed cases live in tests/); prints the worst relative difference and fails above 1e-11."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pxmcmc_b200 import device as D
from pxmcmc_b200.forward import SphericalWaveletTransformOperator
from pxmcmc_b200.mcmc import MYULA, PxMCMCParams
from pxmcmc_b200.prior import S2_Wavelets_L1

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
rel = lambda a, b: float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300))
cases = [(2, 2.0, 0, 1), (3, 2.0, 1, 2), (4, 3.0, 0, 3), (5, 1.5, 2, 2), (33, 2.0, 2, 4), (40, 1.3, 3, 1), (17, 4.0, 1, 5), (64, 2.0, 2, 2),
         (65, 1.5, 2, 3), (129, 1.5, 2, 2)]
for _ in range(16):
    L = int(rng.integers(2, 72))
    B = float(rng.choice([1.3, 1.5, 2.0, 2.5, 3.0]))
    jm = int(np.ceil(np.log(L) / np.log(B)))
    cases.append((L, B, int(rng.integers(0, max(jm, 1))), int(rng.integers(1, 7))))
worst = 0.0
for L, B, J, nch in cases:
    npix = L * (2 * L - 1)
    prm = PxMCMCParams(nsamples=1, nburn=0, ngap=1, delta=1e-3, lmda=5e-3, mu=2.0, verbosity=0, track=[])
    for kind in ("scalar", "per_ring", "real_pairs", "real_pairs_ring"):
        try:
            data = rng.standard_normal(npix) + (0 if kind.startswith("real_pairs") else 1j * rng.standard_normal(npix))
            if kind in ("per_ring", "real_pairs_ring"):
                sig = np.repeat(0.2 + rng.random(L), 2 * L - 1)
            else:
                sig = 0.4
            nc = nch + (nch % 2) if kind.startswith("real_pairs") else nch
            op = SphericalWaveletTransformOperator(data, sig, "synthesis", L, B, J, nchains=nc)
            reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, prm.lmda * prm.mu, L=L, B=B, J_min=J)
        except Exception as e:  # noqa: BLE001  (tilings with an empty scale: the reference's _multires_bandlimits fails the same way)
            print(f"L={L} B={B} J_min={J}: rejected ({type(e).__name__}: {e})")
            break
        X = rng.laplace(size=(nc, op.nparams)) * 0.05
        if not kind.startswith("real_pairs"):
            X = X + 1j * rng.laplace(size=X.shape) * 0.05
        res = {}
        for mode in ("pixels", "carried") + (("pairs",) if kind.startswith("real_pairs") else ()):
            m = MYULA(op, reg, prm, noise="device", nchains=nc, seed=11, stream0=2, real_pairs=(mode == "pairs"))
            op.fuse_ring = mode != "pixels"
            eng = m.engine
            x = eng.pack(X) if mode == "pairs" else m._state(X.astype(complex))
            eng.forward.fuse_ring = mode != "pixels"
            p = eng._initial_preds(x)
            for _ in range(3):
                x, p = eng.iterate(x, p)
            x, p = (m.unpack(x), m.unpack(eng._pix(p))) if mode == "pairs" else (x, eng._pix(p))
            res[mode] = (D.to_host(x), D.to_host(p), getattr(eng.forward, "_ring_kind", lambda: None)() if mode != "pixels" else None)
        op.fuse_ring = True
        for mode in res:
            if mode == "pixels":
                continue
            e = max(rel(res[mode][0], res["pixels"][0]), rel(res[mode][1], res["pixels"][1]))
            worst = max(worst, e)
            flag = "" if e < 1e-11 else "   <-- FAIL"
            print(f"L={L:3d} B={B} J_min={J} chains={nc} {kind:15s} {mode:8s} ({res[mode][2]}): {e:.2e}{flag}")
print(f"worst {worst:.2e}")
sys.exit(0 if worst < 1e-11 else 1)
