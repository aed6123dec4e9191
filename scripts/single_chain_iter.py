"""a few eager single-chain MYULA iterations at L=256 (profiling target)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pxmcmc_b200 import device as D, sht
from pxmcmc_b200.forward import SphericalWaveletTransformOperator
from pxmcmc_b200.mcmc import MYULA, PxMCMCParams
from pxmcmc_b200.prior import S2_Wavelets_L1
L, B, nch = int(os.environ.get("SC_L", 256)), float(os.environ.get("SC_B", 1.5)), int(os.environ.get("SC_NCH", 1))
data = sht.inverse(bench.synthetic_flm(L), L).ravel()
op = SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, 2, nchains=nch)
prm = PxMCMCParams(nsamples=1, delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0, track=[])
reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6, L=L, B=B, J_min=2)
m = MYULA(op, reg, prm, noise="device", nchains=nch)
X = D.to_dev_c(np.random.default_rng(0).laplace(size=(nch, op.nparams)))
P = D.to_dev_c(op.forward(X))
for _ in range(int(os.environ.get("SC_ITERS", 6))):
    X, P = m.iterate(X, P)
torch.cuda.synchronize()
print("ok", float(X.abs().max()))
