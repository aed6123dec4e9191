"""per-stage device time of ONE MYULA chain (eager launches, library events) at L=256 B=1.5 and the table bytes it streams"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pxmcmc_b200 import _lib, device as D, sht
from pxmcmc_b200.forward import SphericalWaveletTransformOperator
from pxmcmc_b200.mcmc import MYULA, PxMCMCParams
from pxmcmc_b200.prior import S2_Wavelets_L1

for L, B, nch in ((256, 1.5, 1), (256, 1.5, 2), (256, 1.5, 4), (512, 2.0, 1)):
    data = sht.inverse(bench.synthetic_flm(L), L).ravel()
    op = SphericalWaveletTransformOperator(data, 1.0, "synthesis", L, B, 2, nchains=nch)
    prm = PxMCMCParams(nsamples=1, delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0, track=[])
    reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6, L=L, B=B, J_min=2)
    m = MYULA(op, reg, prm, noise="device", nchains=nch)
    X = D.to_dev_c(np.random.default_rng(0).laplace(size=(nch, op.nparams)))
    P = D.to_dev_c(op.forward(X))
    for _ in range(50):
        X, P = m.iterate(X, P)
    torch.cuda.synchronize()
    n = 200
    _lib.check(_lib.lib.pxm_profile_begin(16 * n + 64))
    for _ in range(n):
        X, P = m.iterate(X, P)
    ms, cnt = (C.c_double * 3)(), (C.c_longlong * 3)()
    _lib.check(_lib.lib.pxm_profile_end(ms, cnt))
    fam = (C.c_longlong * 4)()
    _lib.check(_lib.lib.pxm_wav_plan_table_bytes_by_family(op.transform._plan(nch).h, fam))
    syn = fam[0] + fam[1]
    print(f"L={L} B={B} chains={nch}: legendre {ms[0]/n*1e3:.1f} us ({cnt[0]//n} launches), ring FFT {ms[1]/n*1e3:.1f} us, elementwise {ms[2]/n*1e3:.1f} us; "
          f"tables of the synthesis pair {syn/1e6:.1f} MB (Lambda_L {fam[0]/1e6:.1f} + W_j {fam[1]/1e6:.1f}), streamed twice per iteration -> "
          f"{2*syn/(ms[0]/n/1e3)/1e9:.0f} GB/s", flush=True)
    del m, op, reg, X, P
    D.WaveletPlan._cache.clear()
    torch.cuda.empty_cache()
