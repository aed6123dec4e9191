"""GPU tests of the sampler machinery added in round 2 (run on the B200 box: `pytest -m gpu`): PxMALA runs longer than
its device-side trace rings, checkpoint / resume of every sampler, CUDA graphs that advance the chain in place, user
subclasses of the prior, complex SKROCK noise, quantiles of long chains."""
import numpy as np
import pytest
from scipy import sparse

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def px():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pxmcmc_b200 import forward, mcmc, measurements, prior, sht, transforms, uncertainty, utils

    class NS:
        pass

    ns = NS()
    ns.forward, ns.mcmc, ns.measurements, ns.prior, ns.sht, ns.transforms, ns.utils, ns.uncertainty = (
        forward, mcmc, measurements, prior, sht, transforms, utils, uncertainty)
    return ns


def _pxmala(px, nsamples, nburn, ngap, noise="device", L=10, nchains=1, track=("logposterior", "chain")):
    B, J = 2.0, 2
    rng = np.random.default_rng(41)
    data = rng.standard_normal(L * (2 * L - 1)) + 0j
    op = px.forward.SphericalWaveletTransformOperator(data, 0.05, "analysis", L, B, J, nchains=nchains)
    p = px.mcmc.PxMCMCParams(nsamples=nsamples, nburn=nburn, ngap=ngap, delta=2e-3, lmda=1e-2, mu=2.0, verbosity=0,
                             track=list(track))
    reg = px.prior.L1("analysis", op.transform.inverse, op.transform.inverse_adjoint, p.lmda * p.mu)
    return px.mcmc.PxMALA(op, reg, p, tune_delta=True, noise=noise, seed=77, stream0=2, nchains=nchains)


def test_pxmala_trace_rings_wrap_and_drain(px):
    """the device-resident loop keeps its acceptance / step-size traces in rings that are drained to the host when
    full: a run with 37-slot rings (several wraps, a partial last ring) gives the traces of a run that never wraps"""
    start = np.random.default_rng(3).laplace(size=190) * 0.1
    a = _pxmala(px, 6, 150, 5)
    a.run(start)
    b = _pxmala(px, 6, 150, 5)
    b.trace_ring = 37
    b.run(start)
    assert len(a.acceptance_trace) > 4 * 37
    assert a.acceptance_trace == b.acceptance_trace and a.deltas_trace == b.deltas_trace
    assert len(a.deltas_trace) == len(a.acceptance_trace) + 1
    assert np.array_equal(a.chain, b.chain) and np.array_equal(a.logPi, b.logPi)


def test_pxmala_seventy_thousand_iterations(px):
    """more iterations than one trace ring holds (65 536): the reference's drivers run nburn = 10^7
    (experiments/weaklensing/main.py:110-119); round 1 aborted here"""
    m = _pxmala(px, 2, 70000, 1, track=("logposterior",))
    assert m._device_resident() and m.trace_ring == 1 << 16
    m.run(np.random.default_rng(3).laplace(size=190) * 0.1)
    n = len(m.acceptance_trace)
    assert n > 70000 and len(m.deltas_trace) == n + 1
    assert 0.05 < np.mean(m.acceptance_trace) < 0.99
    assert set(m.acceptance_trace) == {0, 1} and all(np.isfinite(m.deltas_trace)) and np.isfinite(m.logPi).all()
    # the tuned step stays in the reference's clip interval (mcmc.py:277-279)
    assert min(m.deltas_trace) >= m.lmda * 1e-8 and max(m.deltas_trace) <= m.lmda / 2


@pytest.mark.parametrize("noise", ["device", "host"])
def test_pxmala_checkpoint_resume_continues_the_same_chain(px, noise, tmp_path):
    """both PxMALA loops (device-resident with Philox noise; host-synchronised with numpy's RNG): a run stopped at a
    checkpoint and resumed by a NEW sampler object reproduces the uninterrupted run bit for bit -- state, traces,
    step size, tracked samples"""
    start = np.random.default_rng(3).laplace(size=190) * 0.1
    np.random.seed(5)
    full = _pxmala(px, 8, 6, 3, noise=noise)
    full.trace_ring = 16
    full.run(start)
    np.random.seed(5)
    first = _pxmala(px, 8, 6, 3, noise=noise)
    first.trace_ring = 16
    first.nsamples = 4
    ck = str(tmp_path / "pxmala.npz")
    first.run(start, checkpoint=ck, checkpoint_every=7)
    np.random.seed(999)
    second = _pxmala(px, 8, 6, 3, noise=noise)
    second.trace_ring = 16
    second.run(resume=ck)
    assert list(second.acceptance_trace) == list(full.acceptance_trace)
    assert list(second.deltas_trace) == list(full.deltas_trace)
    assert np.array_equal(second.chain, full.chain) and np.array_equal(second.logPi, full.logPi)
    assert second.delta == full.delta
    with pytest.raises(ValueError):
        px.mcmc.MYULA(full.forward, full.prior, px.mcmc.PxMCMCParams(nsamples=1, verbosity=0), noise=noise).run(resume=ck)


@pytest.mark.parametrize("noise", ["device", "host"])
def test_skrock_checkpoint_resume_continues_the_same_chain(px, noise, tmp_path):
    L, B, J, s_ = 12, 2.0, 2, 3
    rng = np.random.default_rng(8)
    npix = L * (2 * L - 1)
    A = sparse.random(40, npix, density=0.05, random_state=3, format="csr")
    y = rng.standard_normal(40)

    def make(nsamples):
        prm = px.mcmc.PxMCMCParams(delta=1e-7, lmda=5e-8, mu=1.0, s=s_, verbosity=0, nsamples=nsamples, nburn=1, ngap=2,
                                   track=["logposterior", "L2", "prior", "chain", "predictions"])
        op = px.forward.PathIntegralOperator(A, y, 0.1, "synthesis", L, B, J)
        reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 5e-8, L=L, B=B, J_min=J)
        return px.mcmc.SKROCK(op, reg, prm, noise=noise, seed=11, stream0=1)

    start = rng.laplace(size=make(1).forward.nparams) * 0.01
    np.random.seed(3)
    full = make(5)
    full.run(start)
    np.random.seed(3)
    first = make(5)
    first.nsamples = 2
    ck = str(tmp_path / "skrock.npz")
    first.run(start, checkpoint=ck, checkpoint_every=2)
    np.random.seed(4242)
    second = make(5)
    second.run(resume=ck)
    assert np.array_equal(second.chain, full.chain) and np.array_equal(second.logPi, full.logPi)
    assert np.array_equal(second.preds, full.preds)


def test_graphed_chains_advance_in_place(px):
    """CUDA-graph replays write the new state and predictions straight into the buffers they started from (update
    kernel and the operator's last kernel): same chain as the eager loop, no copy-back kernels in the graph"""
    import torch

    from pxmcmc_b200 import device as D

    L, B, J, nch = 24, 1.5, 2, 3
    rng = np.random.default_rng(10)
    data = rng.standard_normal(L * (2 * L - 1)) + 1j * rng.standard_normal(L * (2 * L - 1))
    op = px.forward.SphericalWaveletTransformOperator(data, 0.3, "synthesis", L, B, J, nchains=nch)
    prm = px.mcmc.PxMCMCParams(delta=1e-5, lmda=2e-5, mu=1.0, verbosity=0, nsamples=1, track=[])
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 2e-5, L=L, B=B, J_min=J)
    X0 = D.to_dev_c(rng.laplace(size=(nch, op.nparams)))
    P0 = D.to_dev_c(op.forward(X0))
    eager = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nch, seed=3)
    graphed = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nch, seed=3)
    chain = graphed.capture(X0, graphed._initial_preds(X0), iterations=2)
    ptr_of = lambda P: (P.t if hasattr(P, "t") else P).data_ptr()  # noqa: E731  (ring-carried predictions wrap their tensor)
    ptrs = (chain.X.data_ptr(), ptr_of(chain.P))
    x, p = X0, eager._initial_preds(X0)  # the form run() carries the predictions in (ring coefficients here)
    for _ in range(3):
        for _ in range(2):
            x, p = eager.iterate(x, p)
        chain.step()
        assert torch.equal(chain.state()[0], x) and torch.equal(chain.state()[1], eager._pix(p))
    assert (chain.X.data_ptr(), ptr_of(chain.P)) == ptrs
    # kernels of one replay: only libpxmcmc_b200 kernels (no ATen copy / fill kernels)
    try:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            chain.step()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
    except Exception as exc:  # noqa: BLE001  (no CUPTI on the box: the equality above still holds)
        pytest.skip(f"kernel names unavailable: {exc}")
    kernels = [n for n in names if "memcpy" not in n.lower() and "memset" not in n.lower()]
    assert kernels, names
    assert not [n for n in names if "at::" in n or "memcpy" in n.lower()], names


def test_output_placement_falls_back_to_a_copy(px):
    """`_forward_dev(out=...)` with an operator whose result does not come from the library's allocator (here: a user
    measurement in numpy) still leaves the predictions in `out`"""
    from pxmcmc_b200 import device as D

    class Twice(px.measurements.Measurement):
        def forward(self, X):
            return 2 * X

        def adjoint(self, Y):
            return 2 * Y

    n = 50
    op = px.forward.ForwardOperator(np.zeros(n), 1.0, "analysis", transform=px.transforms.IdentityTransform(),
                                    measurement=Twice(n, n), nparams=n)
    m = px.mcmc.MYULA(op, px.prior.L1("analysis", lambda v: v, lambda v: v, 0.1),
                      px.mcmc.PxMCMCParams(nsamples=1, verbosity=0, track=[]))
    X = D.to_dev_c(np.arange(n, dtype=float)[None, :])
    out = D.to_dev_c(np.zeros((1, n)))
    r = m._forward_dev(X, out=out)
    assert r.data_ptr() == out.data_ptr() and np.array_equal(out.cpu().numpy()[0], 2.0 * np.arange(n))


def test_user_prior_subclass_is_not_bypassed(px):
    """a subclass of L1 that overrides `proxf` / `prior` (the reference's extension pattern) is called through its own
    methods: the fused soft threshold and the device reduction are only taken for the library's implementations"""
    from pxmcmc_b200.prior import is_library_l1

    L, B, J = 10, 2.0, 2
    rng = np.random.default_rng(6)
    data = rng.standard_normal(L * (2 * L - 1)) + 0j
    op = px.forward.SphericalWaveletTransformOperator(data, 0.5, "synthesis", L, B, J)

    class HalfProx(px.prior.L1):
        def proxf(self, X):
            return 0.5 * np.asarray(X)

        def prior(self, X):
            return 3.0 * float(np.sum(np.abs(X) ** 2))

    prm = px.mcmc.PxMCMCParams(delta=1e-3, lmda=2e-3, mu=1.5, verbosity=0, nsamples=1, track=[])
    lib_prior = px.prior.L1("synthesis", None, None, 0.1)
    mine = HalfProx("synthesis", None, None, 0.1)
    assert is_library_l1(lib_prior) and not is_library_l1(mine)
    assert is_library_l1(px.prior.S2_Wavelets_L1_Power_Weights("synthesis", None, None, 0.1, L=L, B=B, J_min=J))
    m = px.mcmc.MYULA(op, mine, prm)
    assert m._fused_prox() is None and not m._native()
    X = rng.laplace(size=op.nparams)
    P = op.forward(X)
    np.random.seed(2)
    w = np.random.randn(op.nparams)
    np.random.seed(2)
    Xn, _ = m.iterate(m._state(X), m._state(P))
    expect = (1 - 0.5) * X + 0.5 * (0.5 * X) - 1e-3 * op.calc_gradg(P) + np.sqrt(2e-3) * w
    assert rel_l2(Xn.cpu().numpy()[0], expect) < 1e-13
    lp, l2, pr = m.logpi(X, P)
    assert np.isclose(pr, 3.0 * np.sum(np.abs(X) ** 2), rtol=1e-13) and np.isclose(lp, -1.5 * pr - l2, rtol=1e-13)
    # PxMALA with such a prior runs the host-synchronised loop
    pm = px.mcmc.PxMALA(op, mine, prm, noise="device")
    assert not pm._device_resident()


def test_skrock_complex_noise(px):
    """`complex=True`: Z = randn + 1j randn drawn real part first (pxmcmc/mcmc.py:344-347); host-noise step against the
    oracle recursion, and the device-noise step runs"""
    from oracle import pxmcmc_ref as R

    L, B, J, s_ = 10, 2.0, 2, 3
    rng = np.random.default_rng(12)
    data = rng.standard_normal(L * (2 * L - 1)) + 1j * rng.standard_normal(L * (2 * L - 1))
    op = px.forward.SphericalWaveletTransformOperator(data, 0.2, "synthesis", L, B, J)
    prm = px.mcmc.PxMCMCParams(delta=1e-6, lmda=5e-7, mu=1.0, s=s_, complex=True, verbosity=0, nsamples=1, track=[])
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 5e-7, L=L, B=B, J_min=J)
    m = px.mcmc.SKROCK(op, reg, prm)
    X = rng.laplace(size=op.nparams) + 1j * rng.laplace(size=op.nparams)
    np.random.seed(4)
    Z = np.random.randn(op.nparams) + 1j * np.random.randn(op.nparams)
    np.random.seed(4)
    Xn = m.chain_step(X)
    t = R.WaveletTransform(L, B, J)
    oop = R.ForwardOperator(data, 0.2, "synthesis", t, R.IdentityMeasurement(data.size, data.size), t.ncoefs)
    oprior = R.S2WaveletsL1("synthesis", t.inverse, t.inverse_adjoint, 5e-7, L, B, J)
    assert rel_l2(Xn, R.skrock_step(oop, oprior, 1e-6, 5e-7, s_, X, Z)) < TOL
    md = px.mcmc.SKROCK(op, reg, prm, noise="device", seed=1)
    out = md.chain_step(X)
    assert np.isfinite(out).all() and np.abs(out.imag - X.imag).max() > 0


def test_pxmala_transition_sums_are_reduced_over_ranks_before_squaring(px):
    """m-sharded operators: every rank holds a partial sum of (X2 - X1 - (delta/2) grad log pi)^2; the reference squares
    the TOTAL again (mcmc.py:286-289), and the accept test must use one uniform on all ranks"""
    import torch

    m = _pxmala(px, 1, 0, 1, noise="host")
    rng = np.random.default_rng(2)
    X1, X2, pr, g = (rng.standard_normal(190) + 1j * rng.standard_normal(190) for _ in range(4))
    plain = m.calc_logtransition(X1, X2, pr, g)
    m.forward._pxm_allreduce = lambda t: 2 * t  # two ranks holding identical halves
    try:
        assert np.isclose(m.calc_logtransition(X1, X2, pr, g), 4 * plain, rtol=1e-14)
        np.random.seed(1)
        u = np.random.rand()
        np.random.seed(1)
        assert m._shared_uniform() == 2 * u  # the (fake) reduction was applied to the draw
        assert not torch.distributed.is_initialized()
    finally:
        del m.forward._pxm_allreduce


def test_credible_interval_of_a_long_chain(px):
    """more samples than the shared-memory sort holds (16 384): same values as numpy's quantile"""
    rng = np.random.default_rng(5)
    chain = rng.standard_normal((20000, 37))
    ours = px.uncertainty.credible_interval_range(chain, alpha=0.1)
    ref = np.quantile(chain, 0.95, axis=0) - np.quantile(chain, 0.05, axis=0)
    assert np.array_equal(ours, ref)
    import torch

    ours_d = px.uncertainty.credible_interval_range(torch.from_numpy(chain).cuda(), alpha=0.1)
    assert np.allclose(ours_d.cpu().numpy(), ref, rtol=1e-14, atol=0)


@pytest.mark.parametrize("noise,nchains", [("device", 1), ("device", 3), ("host", 1)])
def test_async_spill_and_memmapped_chain_equal_the_synchronous_tracking(px, noise, nchains, tmp_path):
    """tracked samples leave the device through pinned staging buffers on a copy stream while the (in-place, graphed)
    chain keeps iterating; `spill_dir` keeps `chain` / `preds` in .npy memory maps: same arrays as the synchronous copy"""
    L, B, J = 12, 2.0, 2
    rng = np.random.default_rng(9)
    data = rng.standard_normal(L * (2 * L - 1)) + 0j
    track = ["logposterior", "L2", "prior", "chain", "predictions"]

    def run(**kw):
        op = px.forward.SphericalWaveletTransformOperator(data, 0.2, "synthesis", L, B, J, nchains=nchains)
        p = px.mcmc.PxMCMCParams(nsamples=7, nburn=3, ngap=2, delta=1e-5, lmda=2e-5, mu=1.0, verbosity=0, track=track)
        reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, p.lmda, L=L, B=B, J_min=J)
        m = px.mcmc.MYULA(op, reg, p, noise=noise, seed=21, stream0=4, nchains=nchains, **kw)
        np.random.seed(31)
        m.run(np.random.default_rng(1).laplace(size=op.nparams))
        return m

    sync = run(async_spill=False)
    spill = run(async_spill=True)
    mm = run(async_spill=True, spill_dir=str(tmp_path / "spill"))
    assert getattr(spill, "_spill", None) is not None and getattr(sync, "_spill", None) is None
    for name in ("chain", "preds", "logPi", "L2s", "priors"):
        assert np.array_equal(getattr(spill, name), getattr(sync, name)), name
        assert np.array_equal(np.asarray(getattr(mm, name)), getattr(sync, name)), name
    assert isinstance(mm.chain, np.memmap) and isinstance(mm.preds, np.memmap)
    mm.chain.flush()
    lead = (nchains,) if nchains > 1 else ()
    on_disk = np.load(str(tmp_path / "spill" / "chain.npy"), mmap_mode="r")
    assert on_disk.shape == lead + (7, sync.forward.nparams) and np.array_equal(on_disk, sync.chain)
    # saving works on memory-mapped tracked arrays too
    from pxmcmc_b200.saving import load_mcmc, save_mcmc

    path = save_mcmc(mm, px.mcmc.PxMCMCParams(nsamples=7), str(tmp_path), "out")
    assert np.array_equal(load_mcmc(path)[0]["chain"], sync.chain)


def _ring_case(px, sig, nchains, L=24, B=1.5, J=2, complex_data=True, **kw):
    rng = np.random.default_rng(14)
    data = rng.standard_normal(L * (2 * L - 1)) + (1j * rng.standard_normal(L * (2 * L - 1)) if complex_data else 0j)
    op = px.forward.SphericalWaveletTransformOperator(data, sig, "synthesis", L, B, J, nchains=nchains)
    prm = px.mcmc.PxMCMCParams(delta=1e-5, lmda=2e-5, mu=1.0, verbosity=0, nsamples=4, nburn=2, ngap=3,
                               track=["logposterior", "L2", "prior", "chain", "predictions"])
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 2e-5, L=L, B=B, J_min=J)
    return op, reg, prm


@pytest.mark.parametrize("sig_kind,nchains", [("scalar", 1), ("scalar", 3), ("scalar_ring", 2), ("per_ring", 2), ("per_ring_ring", 2)])
def test_ring_carried_predictions_equal_the_pixel_composition(px, sig_kind, nchains):
    """Identity measurement behind a wavelet synthesis: the ring FFT that ends Psi and the one that starts the next
    gradient cancel when the inverse covariance is constant along rings; the samplers then carry the predictions as ring
    coefficients.  Same iterations (1e-12), same tracked arrays of run(), same CUDA-graph chain as the pixel composition."""
    import torch

    from pxmcmc_b200 import device as D
    from pxmcmc_b200.forward import RingPreds

    L = 24
    if sig_kind.startswith("scalar"):
        sig = 0.3
    else:  # the reference's per-ring noise level sqrt(sigma^2 / pixel area) (experiments/earthtopography/main.py:92-94)
        sig = np.sqrt(0.05 / px.utils.calc_pixel_areas(L)).flatten()
    op, reg, prm = _ring_case(px, sig, nchains)
    op.fuse_gram = sig_kind not in ("scalar_ring", "per_ring_ring")
    assert op._ring_fusable()
    # harmonic (Gram) form, predictions carried as f_lm: one constant inverse covariance, or -- with per-ring weights in
    # the Gram table -- constants along rings that share a complex phase; `fuse_gram = False`: ring form
    assert op._ring_kind() == {"scalar": "harm", "scalar_ring": "ring", "per_ring": "harm", "per_ring_ring": "ring"}[sig_kind]
    assert (op._gram_weights() is None) == sig_kind.startswith("scalar")
    rng = np.random.default_rng(3)
    X0 = D.to_dev_c(rng.laplace(size=(nchains, op.nparams)))
    ring = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nchains, seed=5)
    lit = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nchains, seed=5)
    Pr = ring._initial_preds(X0)
    assert isinstance(Pr, RingPreds) and ring._ring_mode() and Pr.kind == op._ring_kind()
    Pl = D.to_dev_c(op.forward(X0))
    assert rel_l2(Pr.pixels().cpu().numpy(), Pl.cpu().numpy()) < 1e-12
    back = op.pixels_to_ring(Pl)                                   # the conversion resume / capture use
    assert rel_l2(back.pixels().cpu().numpy(), Pl.cpu().numpy()) < 1e-12
    Xr, Xl = X0, X0
    for _ in range(4):
        Xr, Pr = ring.iterate(Xr, Pr)
        Xl, Pl = lit.iterate(Xl, Pl)
        assert isinstance(Pr, RingPreds)
        assert rel_l2(Xr.cpu().numpy(), Xl.cpu().numpy()) < 1e-12
        assert rel_l2(Pr.pixels().cpu().numpy(), Pl.cpu().numpy()) < 1e-12
    assert rel_l2(op.gradg_from_ring(Pr).cpu().numpy(), op.calc_gradg(Pl).cpu().numpy()) < 1e-12
    # run(): graphed, ring-carried vs the literal composition
    start = rng.laplace(size=op.nparams)
    a = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nchains, seed=9)
    a.run(start)
    op.fuse_ring = False
    try:
        assert not op._ring_fusable()
        b = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nchains, seed=9)
        b.run(start)
    finally:
        op.fuse_ring = True
    for name in ("chain", "preds", "logPi", "L2s", "priors"):
        assert rel_l2(getattr(a, name), getattr(b, name)) < 1e-10, name
    assert rel_l2(a._final_state[1].cpu().numpy(), b._final_state[1].cpu().numpy()) < 1e-10
    # a captured chain started from pixel predictions converts them and advances in place
    g = px.mcmc.MYULA(op, reg, prm, noise="device", nchains=nchains, seed=5)
    chain = g.capture(X0, D.to_dev_c(op.forward(X0)), iterations=2)
    chain.step()
    chain.step()
    assert isinstance(chain.P, RingPreds)
    assert rel_l2(chain.state()[0].cpu().numpy(), Xl.cpu().numpy()) < 1e-12
    assert rel_l2(chain.state()[1].cpu().numpy(), Pl.cpu().numpy()) < 1e-12
    assert torch.isfinite(torch.view_as_real(chain.state()[1])).all()


def test_ring_mode_is_refused_when_it_does_not_apply(px):
    """an inverse covariance that varies along a ring, an analysis setting, a path-integral measurement, a user subclass
    of the transform: the predictions stay pixels"""
    from scipy import sparse as sp

    L = 12
    rng = np.random.default_rng(2)
    sig = 0.1 + rng.random(L * (2 * L - 1))
    op, reg, prm = _ring_case(px, sig, 1, L=L, B=2.0)
    assert not op._ring_fusable() and not px.mcmc.MYULA(op, reg, prm)._ring_mode()
    data = rng.standard_normal(L * (2 * L - 1))
    assert not px.forward.SphericalWaveletTransformOperator(data, 0.1, "analysis", L, 2.0, 2)._ring_fusable()
    A = sp.random(20, L * (2 * L - 1), density=0.1, random_state=1, format="csr")
    assert not px.forward.PathIntegralOperator(A, rng.standard_normal(20), 0.1, "synthesis", L, 2.0, 2)._ring_fusable()

    class Mine(px.transforms.SphericalWaveletTransform):
        def inverse(self, X):
            return 2 * super().inverse(X)

    op2 = px.forward.ForwardOperator(data, 0.1, "synthesis", transform=Mine(L, 2.0, 2),
                                     measurement=px.measurements.Identity(data.size, data.size), nparams=Mine(L, 2.0, 2).ncoefs)
    assert not op2._ring_fusable()
    op3, _, _ = _ring_case(px, 0.1, 1, L=L, B=2.0)
    assert op3._ring_fusable()

    class MineOp(px.forward.SphericalWaveletTransformOperator):  # a user operator with its own gradient is not bypassed
        def calc_gradg(self, preds):
            return 3 * super().calc_gradg(preds)

    assert not MineOp(data, 0.1, "synthesis", L, 2.0, 2)._ring_fusable()
    with pytest.raises(ValueError):
        MineOp(data, 0.1, "synthesis", L, 2.0, 2).paired()


# ------------------------------------------------------------------ real chain pairs
def _pairs_case(px, nchains, real_pairs, noise="device", L=20, B=1.5, J=2, sig=0.3, nsamples=3, nburn=4, ngap=3,
                track=("logposterior", "L2", "prior", "chain", "predictions")):
    rng = np.random.default_rng(12)
    flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
    data = np.ascontiguousarray(px.sht.inverse(flm, L).ravel().real)  # a real field, as the drivers' data are
    op = px.forward.SphericalWaveletTransformOperator(data, sig, "synthesis", L, B, J, nchains=nchains)
    p = px.mcmc.PxMCMCParams(nsamples=nsamples, nburn=nburn, ngap=ngap, delta=1e-4, lmda=2e-3, mu=30.0, verbosity=0,
                             track=list(track))
    reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, p.lmda * p.mu, L=L, B=B, J_min=J)
    return px.mcmc.MYULA(op, reg, p, noise=noise, nchains=nchains, seed=5, stream0=3, real_pairs=real_pairs), op


@pytest.mark.parametrize("noise", ["device", "host"])
def test_real_chain_pairs_are_the_unpacked_chains(px, noise):
    """MYULA(real_pairs=True): two real chains per complex device chain (every linear operator of the synthesis path
    is complex-linear and real) -- the tracked arrays are those of the ordinary sampler, in which every real chain
    travels as a complex chain with a zero imaginary part as in the reference; same noise streams"""
    nch = 6
    a, op = _pairs_case(px, nch, True, noise)
    b, _ = _pairs_case(px, nch, False, noise)
    start = np.random.default_rng(4).laplace(size=(nch, op.nparams)) * 0.05
    np.random.seed(9)
    a.run(start)
    np.random.seed(9)
    b.run(start)
    assert a.engine is not a and a.engine.nchains == nch // 2 and b.engine is b
    assert a.chain.shape == b.chain.shape == (nch, 3, op.nparams)
    for c in range(nch):
        assert rel_l2(a.chain[c], b.chain[c]) < 1e-12, f"chain {c}"
        assert rel_l2(a.preds[c], b.preds[c]) < 1e-12
    assert np.allclose(a.logPi, b.logPi, rtol=1e-11) and np.allclose(a.L2s, b.L2s, rtol=1e-11)
    assert np.allclose(a.priors, b.priors, rtol=1e-12)


def test_real_chain_pairs_checkpoint_resume(px, tmp_path):
    """a real_pairs run resumed from its checkpoint by a new sampler continues the chains bit for bit"""
    nch = 4
    a, op = _pairs_case(px, nch, True, nsamples=6, nburn=2, ngap=2)
    start = np.random.default_rng(4).laplace(size=(nch, op.nparams)) * 0.05
    a.run(start)
    ck = str(tmp_path / "ck.npz")
    b, _ = _pairs_case(px, nch, True, nsamples=6, nburn=2, ngap=2)
    b.nsamples = 3
    b.run(start, checkpoint=ck)
    c, _ = _pairs_case(px, nch, True, nsamples=6, nburn=2, ngap=2)
    c.run(resume=ck)
    assert np.array_equal(a.chain, c.chain) and np.array_equal(a.logPi, c.logPi)


def test_real_chain_pairs_are_refused_when_they_do_not_apply(px):
    L, B, J = 12, 2.0, 2
    rng = np.random.default_rng(1)
    real = rng.standard_normal(L * (2 * L - 1))
    p = px.mcmc.PxMCMCParams(nsamples=1, nburn=0, ngap=1, delta=1e-4, lmda=1e-3, mu=1.0, verbosity=0, track=[])

    def make(data, nchains=2, params=p, cls=None, setting="synthesis", **kw):
        op = px.forward.SphericalWaveletTransformOperator(data, 0.1, setting, L, B, J, nchains=nchains)
        if setting == "synthesis":
            reg = px.prior.S2_Wavelets_L1("synthesis", op.transform.inverse, op.transform.inverse_adjoint, 1e-3, L=L, B=B, J_min=J)
        else:
            reg = px.prior.L1("analysis", op.transform.inverse, op.transform.inverse_adjoint, 1e-3)
        return (cls or px.mcmc.MYULA)(op, reg, params, noise="device", nchains=nchains, real_pairs=True, **kw)

    make(real)  # fine
    with pytest.raises(ValueError):
        make(real + 0j)  # complex data: the reference's covariance rule makes the noise level complex
    with pytest.raises(ValueError):
        make(real, nchains=3)
    with pytest.raises(ValueError):
        make(real, setting="analysis")
    with pytest.raises(ValueError):
        make(real, cls=px.mcmc.PxMALA)
    pc = px.mcmc.PxMCMCParams(nsamples=1, nburn=0, ngap=1, delta=1e-4, lmda=1e-3, mu=1.0, verbosity=0, track=[], complex=True)
    with pytest.raises(ValueError):
        make(real, params=pc)
    m = make(real, nchains=4)
    with pytest.raises(ValueError):
        m.pack(np.ones((4, 5)) * (1 + 1j))


def test_real_chain_pairs_host_pipeline_does_not_depend_on_the_grouping(px):
    """engine.iterate_host on the packed host state: chain groups flowing through the copy / compute pipeline draw the
    Philox streams of their real chains (stream0 + 2 * packed chain, + 1), whatever the grouping"""
    import torch

    nch = 8
    a, op = _pairs_case(px, nch, True, track=())
    b, _ = _pairs_case(px, nch, True, track=())
    X = np.random.default_rng(8).laplace(size=(nch, op.nparams)) * 0.05
    Xp = a.pack(X)
    Pp = a.engine._pix(a.engine._initial_preds(Xp))
    Xh, Ph = Xp.cpu().pin_memory(), Pp.cpu().pin_memory()
    X1, P1 = a.engine.iterate_host(Xh, Ph, groups=1)
    X2, P2 = b.engine.iterate_host(Xh, Ph, groups=[1, 2, 1])
    X1, P1, X2, P2 = (np.asarray(t.cpu() if hasattr(t, "cpu") else t) for t in (X1, P1, X2, P2))
    assert np.array_equal(X1, X2) and np.array_equal(P1, P2)
    # and they are the chains of the device-resident iteration
    c, _ = _pairs_case(px, nch, True, track=())
    X3, P3 = c.engine.iterate(Xp, c.engine._initial_preds(Xp))
    assert rel_l2(X3.cpu().numpy(), X1) < 1e-12 and rel_l2(c.engine._pix(P3).cpu().numpy(), P1) < 1e-12


def test_host_iteration_state_in_state_out(px):
    """MYULA.iterate_host(X): only the state crosses PCIe; the predictions behind the gradient are recomputed on the
    device -- the same new state as with the predictions passed in, whatever the grouping"""
    op, reg, prm = _ring_case(px, 0.7, 4)
    mk = lambda: px.mcmc.MYULA(op, reg, prm, noise="device", nchains=4, seed=3, stream0=5)
    X = np.random.default_rng(2).laplace(size=(4, op.nparams)) * 1e-3 + 0j
    a, b, c = mk(), mk(), mk()
    P = a._forward_dev(a._state(X)).cpu().numpy()
    X1, P1 = a.iterate_host(X, P)
    X2 = np.asarray(b.iterate_host(X))
    X3, P3 = c.iterate_host(X, None, None, np.empty_like(P), groups=[1, 3])
    assert rel_l2(X2, X1) < 1e-12 and np.array_equal(np.asarray(X3), X2)
    assert rel_l2(np.asarray(P3), P1) < 1e-12
