"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, host-side math of the library (tiling, Wigner recurrence) against the
oracle, host logic of the Python mirror, and the world_size-2 gloo path of the
chain sharding."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pxmcmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pxm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from pxmcmc_b200 import _lib

    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/pxmcmc_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pxmcmc_b200 import _lib
    from pxmcmc_b200.transforms import SphericalWaveletTransform
    from pxmcmc_b200.utils import soft

    with pytest.raises(_lib.PxmError):
        SphericalWaveletTransform(8, 2, 2)
    with pytest.raises(_lib.PxmError):
        soft(np.ones(4), 0.5)
    assert _lib.lib.pxm_init(0) != 0
    assert b"no CUDA device" in _lib.lib.pxm_last_error() or b"cuda" in _lib.lib.pxm_last_error().lower()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pxmcmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports the oracle"


@pytest.mark.parametrize("LBJ", [(10, 2, 2), (32, 1.5, 2), (128, 2, 2), (256, 1.5, 2), (28, 2, 2)])
def test_host_tiling_matches_oracle(LBJ):
    from oracle import s2let_ref
    from pxmcmc_b200 import _lib, utils

    L, B, J = LBJ
    Jm, k0, k = _lib.wavelet_tiling_host(L, B, J)
    kap, kap0 = s2let_ref.tiling_axisym(B, L, J)
    assert Jm == s2let_ref.j_max(L, B)
    assert np.array_equal(k0, kap0) and np.array_equal(k, kap[J:])
    assert list(utils._multires_bandlimits(L, B, J)) == s2let_ref.bandlimits(B, L, J)
    phi, psi = utils.wavelet_tiling(B, L, 1, J, 0)
    rphi, rpsi = s2let_ref.wavelet_tiling(B, L, 1, J, 0)
    assert np.allclose(phi, rphi, rtol=1e-15) and np.allclose(psi, rpsi, rtol=1e-15)


@pytest.mark.parametrize("spin", [0, 2, -2, 1])
def test_host_wigner_recurrence_vs_closed_form(spin):
    from oracle import ssht_ref
    from pxmcmc_b200 import _lib

    L = 9
    th, _ = ssht_ref.sample_positions(L)
    err = 0.0
    for m in range(-(L - 1), L):
        for t in range(L):
            row = _lib.wigner_row_host(L, t, m, spin, L)
            for el in range(max(abs(m), abs(spin)), L):
                ref = (-1.0) ** spin * np.sqrt((2 * el + 1) / (4 * np.pi)) * ssht_ref.wigner_d_explicit(el, m, -spin, th[t])
                err = max(err, abs(row[el] - ref))
    assert err < 1e-13


def test_host_wigner_recurrence_large_L_against_oracle_route():
    """the product's theta-recurrence (with exponent tracking for underflowing seeds) against the
    oracle's independent d(pi/2)-Fourier route at a bandlimit where polar seeds underflow"""
    from oracle import ssht_ref
    from pxmcmc_b200 import _lib

    L = 320
    for (el, m) in ((319, 5), (300, 299), (319, 319), (200, 0), (319, -150)):
        flm = np.zeros(L * L, complex)
        flm[el * el + el + m] = 1
        col = ssht_ref.inverse(flm, L, 0)[:, 0].real
        mine = np.array([_lib.wigner_row_host(L, t, m, 0, L)[el] for t in range(L)])
        assert np.abs(col - mine).max() < 1e-11


def test_utils_host_functions_against_reference_values():
    from oracle import pxmcmc_ref as R
    from pxmcmc_b200 import utils

    # reference tests/test_utils.py:54-82
    assert utils.chebyshev1(3, 5) == 3363 and utils.chebyshev2(3, 5) == 6930 and utils.cheb1der(3, 5) == 5945
    assert utils.chebyshev1(5, 0) == 1 and utils.chebyshev2(2, 1) == 4 and utils.cheb1der(5, 0) == 0
    with pytest.raises(ValueError):
        utils.chebyshev1(1.0, -1)
    assert np.isclose(utils.pixel_area(1, 0, np.pi, 0, 2 * np.pi), 4 * np.pi)
    assert np.isclose(utils.polar_cap_area(1, np.pi / 2), 2 * np.pi)
    assert np.isclose(utils.calc_pixel_areas(10).sum(), 4 * np.pi)
    for L in (3, 10, 33):
        assert np.allclose(utils.mw_map_weights(L), R.mw_map_weights(L), rtol=0, atol=1e-17)
        assert np.isclose(utils.mw_map_weights(L).sum(), 4 * np.pi)
    # layout known answer, reference tests/test_utils.py:8-32
    w = np.ones((861, 9)) + np.arange(9)[None, :]
    assert np.array_equal(utils.flatten_mlm(w, np.zeros(861)), np.concatenate([[i] * 861 for i in range(10)]))
    fw, fs = utils.expand_mlm(np.ones(8610), nscales=9)
    assert fw.shape == (861, 9) and fs.shape == (861,)
    a, b = utils.expand_mlm(np.arange(10.0), nscalcoefs=3)
    assert np.array_equal(b, [0, 1, 2]) and np.array_equal(a, np.arange(3.0, 10))
    with pytest.raises(ValueError):
        utils.expand_mlm(np.ones(4))
    assert np.array_equal(utils.hard(np.arange(1.0, 11), 0.3), [0, 0, 0, 0, 0, 0, 0, 8, 9, 10])


def test_prior_weight_vectors_and_skrock_coefficients_host():
    """host-side setup of the mirror classes against the fixtures of the unmodified reference"""
    from conftest import golden
    from pxmcmc_b200.prior import S2_Wavelets_L1, S2_Wavelets_L1_Power_Weights
    from pxmcmc_b200.mcmc import PxMCMCParams

    g = golden("ref_wavelet_L10B2.npz")
    assert np.allclose(S2_Wavelets_L1("synthesis", None, None, 1.0, 10, 2, 2).T, g["s2_T"], rtol=1e-14)
    pw = S2_Wavelets_L1_Power_Weights("synthesis", None, None, 1.0, 10, 2, 2, eta=1)
    assert np.allclose(pw.T, g["s2pw_T"], rtol=1e-14) and np.allclose(pw.map_weights, g["s2pw_w"], rtol=1e-14)
    with pytest.raises(NotImplementedError):
        S2_Wavelets_L1("analysis", None, None, 1.0, 10, 2, 2)
    p = PxMCMCParams()
    assert (p.lmda, p.delta, p.s, p.mu, p.nsamples, p.nburn, p.ngap, p.complex, p.verbosity) == (3e-5, 1e-5, 1, 1, int(1e6), int(1e3), int(1e2), False, 100)


def test_forward_operator_inverse_covariance_rules():
    from pxmcmc_b200.forward import ForwardOperator

    d = np.arange(4.0)
    op = ForwardOperator(d, 0.1, "analysis")
    assert np.allclose(op.invcov.diagonal(), 100.0) and np.allclose(op._diag, 100.0)
    op = ForwardOperator(d + 0j, 0.1, "analysis")  # complex data, real sigma (forward.py:80-82)
    assert np.allclose(op._diag, (1 - 1j) / (np.sqrt(2) * 0.01))
    op = ForwardOperator(d, np.array([1.0, 2, 4, 5]), "synthesis")
    assert np.allclose(op._diag, 1 / np.array([1.0, 4, 16, 25]))
    with pytest.raises(ValueError):
        ForwardOperator(d, 0.1, "other")
    with pytest.raises(ValueError):
        ForwardOperator(d, np.ones((2, 3)), "analysis")
    with pytest.raises(TypeError):
        ForwardOperator(d, np.ones(3), "analysis")


def test_chain_sharding_partition():
    from pxmcmc_b200.sharding import chain_shard

    for total in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [chain_shard(total, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == total
            pos = 0
            for s, c in parts:
                assert s == pos
                pos += c
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        chain_shard(4, 2, 2)


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from pxmcmc_b200.sharding import chain_shard, gather_chains, max_over_ranks, philox_stream0
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world, total = dist.get_rank(), dist.get_world_size(), 5
start, count = chain_shard(total, world, rank)
assert philox_stream0(total, world, rank) == start
local = np.stack([np.full(3, float(start + i)) for i in range(count)]) if count else np.zeros((0, 3))
out = gather_chains(local, total)
m = max_over_ranks(10.0 + rank)
assert m == 10.0 + world - 1
if rank == 0:
    assert out.shape == (total, 3) and np.array_equal(out[:, 0], np.arange(total, dtype=float))
    print("GATHER_OK")
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
"""


def test_chain_sharding_gloo_world_size_2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "GATHER_OK" in outs[0][0]


# ---------------------------------------------------------------- m-sharded partition (host logic)
def test_msharded_partition_is_a_partition():
    """every azimuthal order and every ring of every grid has exactly one owner; ring ranges are
    whole blocks of 64; the snake keeps the triangular Legendre work balanced"""
    from pxmcmc_b200 import msharded as ms

    for world in (1, 2, 3, 4, 8):
        for ell in (3, 64, 65, 130, 256, 512, 1024):
            for rot in (0, 1, 5):
                seen = np.zeros(ell, dtype=int)
                for r in range(world):
                    t0, t1 = ms.ring_range(ell, rot, r, world)
                    assert t0 % 64 == 0 and 0 <= t0 <= t1 <= ell
                    seen[t0:t1] += 1
                assert np.all(seen == 1)
        L = 512
        work = np.zeros(world)
        for am in range(L):
            o = ms.owner_of_m(am, world)
            assert 0 <= o < world
            work[o] += L - am
        assert work.max() / work.mean() < 1.02
    masks = [ms.flm_owner_mask(12, r, 3) for r in range(3)]
    assert np.array_equal(np.sum(masks, axis=0), np.ones(144))
    with pytest.raises(Exception):
        ms.ring_range(16, 0, 2, 2)


def test_msharded_layout_roundtrip():
    from pxmcmc_b200 import msharded as ms

    ells, world = [4, 8, 100, 200, 200], 4
    rng = np.random.default_rng(0)
    n = sum(e * (2 * e - 1) for e in ells)
    full = rng.standard_normal(n)
    back = np.zeros(n)
    total = 0
    for r in range(world):
        lay = ms.MapLayout(ells, range(1, len(ells) + 1), r, world)
        loc = lay.to_local(full)
        assert loc.shape == (lay.n_local,)
        lay.scatter_into(back, loc)
        total += lay.n_local
    assert total == n and np.array_equal(back, full)
    # small grids are rotated over the ranks instead of piling up on rank 0
    owners = {next(r for r in range(world) if ms.ring_range(8, rot, r, world)[1] > 0) for rot in range(4)}
    assert owners == {0, 1, 2, 3}


_MS_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from pxmcmc_b200 import msharded as ms
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
# scalar reductions of a sharded sampler: partial sums per rank -> identical totals everywhere
red = ms.allreduce_sum()
ells = [4, 8, 70, 130]
lay = ms.MapLayout(ells, range(1, 5), rank, world)
full = np.arange(lay.n_full, dtype=float)
part = torch.tensor([lay.to_local(full).sum()], dtype=torch.float64)
tot = red(part)
assert part.item() != tot.item() or world == 1
assert tot.item() == full.sum()
# local -> full reassembly through all_gather_object (what a driver does to save a sharded chain)
pieces = [None] * world
dist.all_gather_object(pieces, (lay.index, lay.to_local(full)))
back = np.full(lay.n_full, -1.0)
for idx, loc in pieces:
    back[idx] = loc
assert np.array_equal(back, full)
if rank == 0:
    print("MS_HOST_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_msharded_host_logic_gloo_world_size_2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "ms_worker.py"
    script.write_text(_MS_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "MS_HOST_OK" in outs[0][0]


def test_quantile_virtual_index_reproduces_numpy_and_the_reference_fixture():
    """host half of credible_interval_range: numpy's virtual index / _lerp arithmetic, checked bit for
    bit against np.quantile and against the vectors the unmodified reference produced
    (tests/golden/ref_uncertainty.npz, oracle/gen_golden_uncertainty.py)"""
    from conftest import golden
    from oracle.gen_golden_uncertainty import CASES, make_chain
    from pxmcmc_b200.uncertainty import _virtual_index

    def lerp(a, b, t):
        d = b - a
        return np.where(t >= 0.5, b - d * (1.0 - t), a + d * t)

    g = golden("ref_uncertainty.npz")
    for i, (n, p, seed) in enumerate(CASES):
        s = np.sort(make_chain(n, p, seed), axis=0)
        for alpha, key in ((0.05, f"ci_{i}"), (0.1, f"ci10_{i}")):
            q = []
            for qq in (alpha / 2, 1 - alpha / 2):
                lo, gam = _virtual_index(n, qq)
                q.append(lerp(s[lo], s[min(lo + 1, n - 1)], gam))
                assert np.array_equal(q[-1], np.quantile(s, qq, axis=0))
            assert np.array_equal(q[1] - q[0], g[key])
    with pytest.raises(ValueError):
        _virtual_index(10, 1.5)


def test_saving_roundtrip_uses_the_reference_dataset_names(tmp_path):
    """save_mcmc / load_mcmc (pxmcmc/saving.py:18-36): dataset and attribute names of the reference"""
    from pxmcmc_b200 import saving
    from pxmcmc_b200.mcmc import PxMCMCParams

    class Run:
        pass

    run = Run()
    run.logPi, run.L2s, run.priors = np.arange(4.0), np.arange(4.0) * 2, np.arange(4.0) * 3
    run.chain, run.preds = np.arange(12.0).reshape(4, 3), np.ones((4, 5))
    run.acceptance_trace, run.deltas_trace = [1, 0, 1, 1], [1e-6, 2e-6, 2e-6, 3e-6, 3e-6]
    prm = PxMCMCParams(nsamples=4, nburn=1, ngap=2, track=["chain"])
    path = saving.save_mcmc(run, prm, str(tmp_path), filename="out", L=10, setting="synthesis")
    data, attrs = saving.load_mcmc(path)
    assert sorted(data) == sorted(["logposterior", "predictions", "chain", "L2s", "priors", "acceptances", "deltas"])
    assert np.array_equal(data["chain"], run.chain) and data["acceptances"].dtype == np.int8
    assert int(attrs["nsamples"]) == 4 and int(attrs["L"]) == 10 and str(attrs["setting"]) == "synthesis"
    for k in prm.__dict__:
        assert k in attrs


def test_dft_codelets_on_the_host(tmp_path):
    """the register DFT codelets of the ring FFT (multiply-add butterflies, pruned half-input / half-output radix-16/32
    transforms) compiled for the HOST by nvcc and compared with a direct long-double DFT: 2e-15 relative"""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "dft_codelets")
    subprocess.run([nvcc, "-Wno-deprecated-gpu-targets", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "pxmcmc_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "native", "dft_codelets.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "worst" in r.stdout
