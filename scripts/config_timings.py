"""Per-iteration timings of the BASELINE.json configurations that are parity-test cases rather
than bench lines (single-chain MYULA / PxMALA / SKROCK / weak lensing), device-resident state."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy import sparse
import bench
from pxmcmc_b200 import device as D, sht, _lib
from pxmcmc_b200.forward import SphericalWaveletTransformOperator, PathIntegralOperator, ForwardOperator
from pxmcmc_b200.measurements import WeakLensing
from pxmcmc_b200.transforms import SphericalWaveletTransform
from pxmcmc_b200.mcmc import MYULA, PxMALA, SKROCK, PxMCMCParams
from pxmcmc_b200.prior import L1, S2_Wavelets_L1, S2_Wavelets_L1_Power_Weights


def timeit(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    l0 = _lib.lib.pxm_launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    return dt, (_lib.lib.pxm_launch_count() - l0) / n


def data_map(L):
    d = sht.inverse(bench.synthetic_flm(L), L).ravel()
    return d / np.sqrt(np.mean(np.abs(d) ** 2))


out = {}
which = sys.argv[1:] or ["1", "1c", "2", "3", "4"]

if "1" in which or "1c" in which:
    for tag, L, B in (("config1 MYULA L=32 B=1.5 synthesis", 32, 1.5), ("single-chain MYULA L=256 B=1.5 synthesis", 256, 1.5)):
        if tag.startswith("config1") and "1" not in which:
            continue
        if tag.startswith("single") and "1c" not in which:
            continue
        op = SphericalWaveletTransformOperator(data_map(L), 1.0, "synthesis", L, B, 2)
        prm = PxMCMCParams(delta=1e-6, lmda=1e-6, mu=1.0, verbosity=0, nsamples=1, track=[])
        reg = S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6, L=L, B=B, J_min=2)
        m = MYULA(op, reg, prm, noise="device")
        st = [D.to_dev_c(np.random.default_rng(0).laplace(size=(1, op.nparams)))]
        st.append(D.to_dev_c(op.forward(st[0])))
        def f():
            st[0], st[1] = m.iterate(st[0], st[1])
        dt, nl = timeit(f, 200)
        out[tag] = {"ms_per_iteration": dt * 1e3, "iterations_per_s": 1 / dt, "launches": nl}
        chain = m.capture(st[0], st[1], iterations=1)
        dtg, _ = timeit(chain.step, 500)
        out[tag].update({"graph_ms_per_iteration": dtg * 1e3, "graph_iterations_per_s": 1 / dtg})
        print(tag, out[tag], flush=True)

if "2" in which:
    L, B = 256, 1.5
    op = SphericalWaveletTransformOperator(data_map(L), 0.1, "analysis", L, B, 2)
    prm = PxMCMCParams(delta=1e-7, lmda=1e-6, mu=1.0, verbosity=0, nsamples=40, nburn=0, ngap=1, track=["logposterior"])
    reg = L1("analysis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6)
    m = PxMALA(op, reg, prm, tune_delta=True)
    np.random.seed(0)
    m.run(np.zeros(op.nparams))  # warm
    m = PxMALA(op, reg, prm, tune_delta=True)
    torch.cuda.synchronize()
    l0 = _lib.lib.pxm_launch_count()
    t0 = time.perf_counter()
    m.run(np.zeros(op.nparams))
    torch.cuda.synchronize()
    nit = len(m.acceptance_trace)
    dt = (time.perf_counter() - t0) / nit
    out["config2 PxMALA L=256 B=1.5 analysis (host accept, host noise)"] = {"ms_per_iteration": dt * 1e3, "iterations_per_s": 1 / dt, "iterations": nit, "launches": (_lib.lib.pxm_launch_count() - l0) / nit, "acceptance": float(np.mean(m.acceptance_trace))}
    prm.nsamples, prm.ngap = 41, 100  # the reference's default thinning: accepted samples stored every 100 iterations
    for tag, dev_loop in (("host accept, Philox noise", False), ("device accept and step-size tuning, Philox noise", True)):
        m = PxMALA(op, reg, prm, tune_delta=True, noise="device", seed=3)
        if not dev_loop:
            m._device_resident = lambda: False
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.run(np.zeros(op.nparams))
        torch.cuda.synchronize()
        nit = len(m.acceptance_trace)
        dt = (time.perf_counter() - t0) / nit
        out[f"config2 PxMALA L=256 B=1.5 analysis ({tag}, ngap=100)"] = {"ms_per_iteration": dt * 1e3, "iterations_per_s": 1 / dt, "iterations": nit, "acceptance": float(np.mean(m.acceptance_trace))}
    print(out, flush=True)

if "2b" in which:
    # 64 PxMALA chains as one batch (analysis prior, L=256): the device-resident loop with per-chain step sizes
    L, B, nch = 256, 1.5, 64
    op = SphericalWaveletTransformOperator(data_map(L), 0.1, "analysis", L, B, 2, nchains=nch)
    prm = PxMCMCParams(delta=1e-7, lmda=1e-6, mu=1.0, verbosity=0, nsamples=3, nburn=0, ngap=100, track=["logposterior"])
    reg = L1("analysis", op.transform.inverse, op.transform.inverse_adjoint, 1e-6)
    m = PxMALA(op, reg, prm, tune_delta=True, noise="device", seed=3, nchains=nch)
    m.run(np.zeros(op.nparams))  # plans, tables, graph capture
    prm.nsamples = 11
    m = PxMALA(op, reg, prm, tune_delta=True, noise="device", seed=3, nchains=nch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.run(np.zeros(op.nparams))
    torch.cuda.synchronize()
    nit = m.acceptance_trace.shape[1]
    dt = (time.perf_counter() - t0) / nit
    out["64 PxMALA chains L=256 B=1.5 analysis (device-resident loop)"] = {"ms_per_step": dt * 1e3, "chain_iterations_per_s": nch / dt, "iterations": nit, "acceptance": float(m.acceptance_trace.mean())}
    print(out, flush=True)

if "3" in which:
    L, B, s = 128, 2, 10
    rng = np.random.default_rng(7)
    npix = L * (2 * L - 1)
    # synthetic great-circle-like CSR: 10^4 paths, ~2.5 L pixels each, rows sum to 1
    nnz_row = int(2.5 * L)
    rows = np.repeat(np.arange(10000), nnz_row)
    cols = rng.integers(0, npix, size=rows.size)
    A = sparse.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(10000, npix))
    A = sparse.diags(1.0 / np.asarray(A.sum(axis=1)).ravel()) @ A
    truth = data_map(L).real
    y = A @ truth + 0.05 * rng.standard_normal(10000)
    op = PathIntegralOperator(A.tocsr(), y, np.full(10000, 0.05), "synthesis", L, B, 2)
    prm = PxMCMCParams(delta=1e-6, lmda=5e-7, mu=1.0, s=s, verbosity=0, nsamples=1, track=[])
    reg = S2_Wavelets_L1_Power_Weights("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 5e-7, L=L, B=B, J_min=2, eta=1)
    m = SKROCK(op, reg, prm)
    st = [D.to_dev_c(np.zeros((1, op.nparams)))]
    def f():
        st[0] = m._chain_step_dev(st[0])
    dt, nl = timeit(f, 20)
    out["config3 SKROCK s=10 PathIntegral 1e4x32640 (nnz %d) L=128 B=2" % A.nnz] = {"ms_per_step": dt * 1e3, "steps_per_s": 1 / dt, "gradient_evaluations_per_step": s, "launches": nl}
    mg = SKROCK(op, reg, prm, noise="device", seed=5)
    gr = mg.capture(st[0])
    dtg, _ = timeit(gr.step, 50)
    out["config3 SKROCK s=10 PathIntegral 1e4x32640 (nnz %d) L=128 B=2" % A.nnz].update({"graph_ms_per_step": dtg * 1e3, "graph_steps_per_s": 1 / dtg})
    print(out, flush=True)

if "4" in which:
    L, B = 512, 2
    th = (2 * np.arange(L) + 1) * np.pi / (2 * L - 1)
    mask = np.ones((L, 2 * L - 1), dtype=int)
    mask[np.abs(90 - np.degrees(th)) < 10, :] = 0
    wl = WeakLensing(L, mask=mask, ngal=np.full((L, 2 * L - 1), 30.0))
    tr = SphericalWaveletTransform(L, B, 2)
    kappa = data_map(L)
    gdata = wl.forward(kappa)
    op = ForwardOperator(gdata, 1 / wl.inv_cov, "synthesis", transform=tr, measurement=wl, nparams=tr.ncoefs)
    prm = PxMCMCParams(delta=5e-7, lmda=1e-6, mu=1.0, verbosity=0, nsamples=1, track=[])
    reg = S2_Wavelets_L1("synthesis", tr.inverse, tr.inverse_adjoint, 1e-6, L=L, B=B, J_min=2)
    m = MYULA(op, reg, prm, noise="device")
    st = [D.to_dev_c(np.zeros((1, op.nparams)))]
    st.append(D.to_dev_c(op.forward(st[0])))
    def f():
        st[0], st[1] = m.iterate(st[0], st[1])
    dt, nl = timeit(f, 20)
    chain = m.capture(st[0], st[1], iterations=1)
    dtg, _ = timeit(chain.step, 50)
    out["config4 MYULA WeakLensing spin-2 L=512 B=2 single chain, 1 GPU"] = {"ms_per_iteration": dt * 1e3, "iterations_per_s": 1 / dt, "launches": nl, "ndata": int(mask.sum()), "graph_ms_per_iteration": dtg * 1e3, "graph_iterations_per_s": 1 / dtg}
    print(out, flush=True)

json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "config_timings.json"), "w"), indent=1)
