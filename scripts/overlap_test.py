"""Experiment: split the chains of one GPU into G groups stepped on G CUDA streams so that the
tensor-pipe-bound Legendre kernels of one group overlap the LSU-bound ring FFTs of another."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pxmcmc_b200 import device as D, sht
from pxmcmc_b200.forward import SphericalWaveletTransformOperator
from pxmcmc_b200.mcmc import MYULA, PxMCMCParams
from pxmcmc_b200.prior import S2_Wavelets_L1

L, B, J = 256, 1.5, 2
total = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = sht.inverse(bench.synthetic_flm(L), L).ravel()
data = data / np.sqrt(np.mean(np.abs(data) ** 2))
for G in (1, 2, 4):
    nch = total // G
    groups = []
    for gi in range(G):
        # distinct plan objects per group (own workspace): bypass the plan cache
        D.WaveletPlan._cache.clear()
        op = SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, J, nchains=nch)
        prm = PxMCMCParams(nsamples=1, nburn=0, ngap=1, delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0, track=[])
        reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6, L=L, B=B, J_min=J)
        m = MYULA(op, reg, prm, noise="device", nchains=nch, seed=1, stream0=gi * nch)
        X = D.to_dev_c(np.random.default_rng(gi).laplace(size=(nch, op.nparams)))
        P = D.to_dev_c(op.forward(X))
        groups.append([m, X, P, torch.cuda.Stream()])
    torch.cuda.synchronize()
    def step():
        for g in groups:
            with torch.cuda.stream(g[3]):
                g[1], g[2] = g[0].iterate(g[1], g[2])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    K = 10
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for g in groups:
        g[3].wait_stream(torch.cuda.current_stream())
    for _ in range(K):
        step()
    for g in groups:
        torch.cuda.current_stream().wait_stream(g[3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"groups={G} chains/group={nch}: {ms:.3f} ms per step -> {total / ms * 1e3:.0f} chain-it/s (host {1e3*(time.perf_counter()-t0)/K:.3f} ms/step)", flush=True)
    del groups
    torch.cuda.empty_cache()
