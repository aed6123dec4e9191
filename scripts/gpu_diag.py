"""First-contact diagnostics on a B200: stage-wise errors of every transform
against the oracle, with the Legendre contraction routed through the plain
kernel and through the DMMA kernel.  Prints; never raises."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import pxmcmc_ref as R  # noqa: E402
from oracle import s2let_ref, ssht_ref  # noqa: E402
from pxmcmc_b200 import _lib, device as D, sht  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def guard(name, fn):
    try:
        t = time.time()
        out = fn()
        torch.cuda.synchronize()
        print(f"[ok ] {name}: {out}  ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    rng = np.random.default_rng(0)

    def elem():
        x = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
        T = np.abs(rng.standard_normal(1000)) * 0.5
        from pxmcmc_b200.utils import soft
        a = soft(x, T)
        b = R.soft(x, T)
        return f"soft complex exact={np.array_equal(a, b)} real exact={np.array_equal(soft(x.real, 0.3), R.soft(x.real, 0.3))}"
    guard("soft", elem)

    def sht_case(L, spin, naive):
        _lib.lib.pxm_debug_set_naive(naive)
        D.ShtPlan._cache.clear()
        flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
        flm[: spin * spin] = 0
        f = rng.standard_normal((L, 2 * L - 1)) + 1j * rng.standard_normal((L, 2 * L - 1))
        res = {}
        res["inv"] = rel(sht.inverse(flm, L, Spin=spin), ssht_ref.inverse(flm, L, spin))
        res["inv_adj"] = rel(sht.inverse_adjoint(f, L, Spin=spin), ssht_ref.inverse_adjoint(f, L, spin))
        res["fwd"] = rel(sht.forward(f, L, Spin=spin), ssht_ref.forward(f, L, spin))
        res["fwd_adj"] = rel(sht.forward_adjoint(flm, L, Spin=spin), ssht_ref.forward_adjoint(flm, L, spin))
        _lib.lib.pxm_debug_set_naive(0)
        return " ".join(f"{k}={v:.2e}" for k, v in res.items())

    for L in (4, 12, 33):
        for spin in (0, 2):
            for naive in (1, 0):
                guard(f"sht L={L} spin={spin} naive={naive}", lambda L=L, spin=spin, naive=naive: sht_case(L, spin, naive))

    def wav_case(L, B, J, naive):
        _lib.lib.pxm_debug_set_naive(naive)
        D.WaveletPlan._cache.clear()
        from pxmcmc_b200.transforms import SphericalWaveletTransform
        t = SphericalWaveletTransform(L, B, J)
        o = R.WaveletTransform(L, B, J)
        xp = rng.standard_normal(L * (2 * L - 1)) + 1j * rng.standard_normal(L * (2 * L - 1))
        xc = rng.standard_normal(t.ncoefs) + 1j * rng.standard_normal(t.ncoefs)
        res = {
            "ncoefs": (t.ncoefs == o.ncoefs),
            "inverse": rel(t.inverse(xc), o.inverse(xc)),
            "inverse_adjoint": rel(t.inverse_adjoint(xp), o.inverse_adjoint(xp)),
            "forward": rel(t.forward(xp), o.forward(xp)),
            "forward_adjoint": rel(t.forward_adjoint(xc), o.forward_adjoint(xc)),
        }
        _lib.lib.pxm_debug_set_naive(0)
        return " ".join(f"{k}={v if isinstance(v, bool) else format(v, '.2e')}" for k, v in res.items())

    for (L, B, J) in ((10, 2, 2), (32, 1.5, 2)):
        for naive in (1, 0):
            guard(f"wavelet L={L} B={B} naive={naive}", lambda L=L, B=B, J=J, naive=naive: wav_case(L, B, J, naive))

    def big(L):
        flm = rng.standard_normal(L * L) + 1j * rng.standard_normal(L * L)
        t0 = time.time()
        f = sht.inverse(flm, L)
        torch.cuda.synchronize()
        t1 = time.time() - t0
        fl2 = sht.forward(f, L)
        e_rt = rel(fl2, flm)
        ref = ssht_ref.inverse(flm, L, 0)
        return f"inverse vs oracle {rel(f, ref):.2e}, forward(inverse) roundtrip {e_rt:.2e}, first-call {t1:.2f}s"
    for L in (128, 256):
        guard(f"sht big L={L}", lambda L=L: big(L))

    def batched():
        L, nb = 16, 5
        flm = rng.standard_normal((nb, L * L)) + 1j * rng.standard_normal((nb, L * L))
        out = D.to_host(D.ShtPlan.get(L, 0, nb).inverse(D.to_dev_c(flm)))
        ref = np.stack([ssht_ref.inverse(flm[i], L, 0).ravel() for i in range(nb)])
        return f"batched inverse {rel(out, ref):.2e}"
    guard("batched", batched)

    def dgemm():
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return f"cuBLAS DGEMM {n}^3: {2*n**3/best/1e9:.1f} TFLOP/s (best of 5, {best:.2f} ms)"
    guard("dgemm peak", dgemm)


if __name__ == "__main__":
    main()
