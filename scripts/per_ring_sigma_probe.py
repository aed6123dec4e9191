import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from pxmcmc_b200 import device as D, sht
from pxmcmc_b200.forward import SphericalWaveletTransformOperator
from pxmcmc_b200.mcmc import MYULA, PxMCMCParams
from pxmcmc_b200.prior import S2_Wavelets_L1
from pxmcmc_b200.utils import calc_pixel_areas
L, B, J, nch = 256, 1.5, 2, 64
data = sht.inverse(bench.synthetic_flm(L), L).ravel()
data = data / np.sqrt(np.mean(np.abs(data) ** 2))
sig = np.sqrt(1e-4 / calc_pixel_areas(L)).flatten()  # the drivers' per-ring noise level (earthtopography/main.py:92-94)
for gram in (True, False):
    op = SphericalWaveletTransformOperator(data, sig, "synthesis", L, B, J, nchains=nch)
    op.fuse_gram = gram
    prm = PxMCMCParams(nsamples=1, nburn=0, ngap=1, delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0, track=[])
    reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6, L=L, B=B, J_min=J)
    m = MYULA(op, reg, prm, noise="device", nchains=nch, seed=1)
    X = D.to_dev_c(np.random.default_rng(0).laplace(size=(nch, op.nparams)))
    P = m._initial_preds(X)
    for _ in range(30):
        X, P = m.iterate(X, P)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        X, P = m.iterate(X, P)
    b.record(); torch.cuda.synchronize()
    print("per-ring sigma, kind", op._ring_kind(), "ms/step", a.elapsed_time(b) / 100, flush=True)
    del m, op, X, P
    bench.release_plans()
