"""iterate_host (the e2e path of bench.py) at the bench shape for several chain-group counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pxmcmc_b200 import device as D, sht
from pxmcmc_b200.forward import SphericalWaveletTransformOperator
from pxmcmc_b200.mcmc import MYULA, PxMCMCParams
from pxmcmc_b200.prior import S2_Wavelets_L1

L, B, J_min, nch = 256, 1.5, 2, 64
data = sht.inverse(bench.synthetic_flm(L), L).ravel()
data = data / np.sqrt(np.mean(np.abs(data) ** 2))
op = SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, J_min, nchains=nch)
prm = PxMCMCParams(nsamples=1, nburn=0, ngap=1, delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0, track=[])
reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6, L=L, B=B, J_min=J_min)
m = MYULA(op, reg, prm, noise="device", nchains=nch, seed=1234)
X = D.to_dev_c(np.random.default_rng(7).laplace(size=(nch, op.nparams)))
P = D.to_dev_c(op.forward(X))
Xh, Ph = X.cpu().pin_memory(), P.cpu().pin_memory()
Xo, Po = torch.empty_like(Xh).pin_memory(), torch.empty_like(Ph).pin_memory()
for groups in [None, 8, [2, 6] + [8] * 7, [1, 3, 4] + [8] * 7, [2, 6] + [8] * 6 + [6, 2]]:
    m.iterate_host(Xh, Ph, Xo, Po, groups=groups)
    m.iterate_host(Xh, Ph, Xo, Po, groups=groups)
    torch.cuda.synchronize()
    t = time.perf_counter()
    n = 5
    for _ in range(n):
        m.iterate_host(Xh, Ph, Xo, Po, groups=groups)
    dt = (time.perf_counter() - t) / n
    print(f"groups {groups}: {dt * 1e3:.2f} ms per step, {nch / dt:.0f} chain-it/s", flush=True)

