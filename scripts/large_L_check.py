"""Adjoint dot-tests and timings of the wavelet operators at a large bandlimit (default L=1024, B=2):
<Psi x, y> = <x, Psi^dagger y> for the synthesis pair and the analysis pair, round trip analysis -> synthesis."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pxmcmc_b200 import device as D

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
t0 = time.perf_counter()
plan = D.WaveletPlan.get(L, B, 2, 1)
g = torch.Generator(device="cuda").manual_seed(3)
rnd = lambda n: torch.randn(n, dtype=torch.float64, device="cuda", generator=g) + 1j * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
x, y = rnd(plan.ncoefs), rnd(plan.npix)
out = {}
for name, fn, a in (("synthesis", plan.synthesis, x), ("synthesis_adjoint", plan.synthesis_adjoint, y),
                    ("analysis", plan.analysis, y), ("analysis_adjoint", plan.analysis_adjoint, x)):
    r = fn(a)
    torch.cuda.synchronize()
    if name == "synthesis":
        print(f"L={L} B={B}: ncoefs {plan.ncoefs}, tables {plan.table_bytes / 2**30:.2f} GiB, first call after {time.perf_counter() - t0:.1f} s", flush=True)
    t = time.perf_counter()
    for _ in range(5):
        r = fn(a)
    torch.cuda.synchronize()
    out[name] = r
    print(f"  {name}: {(time.perf_counter() - t) / 5 * 1e3:.2f} ms", flush=True)
dot = lambda a, b: torch.vdot(a, b)
e1 = abs(dot(y, out["synthesis"]) - dot(out["synthesis_adjoint"], x)) / abs(dot(y, out["synthesis"]))
e2 = abs(dot(x, out["analysis"]) - dot(out["analysis_adjoint"], y)) / abs(dot(x, out["analysis"]))
# exactness: synthesis(analysis(f)) = f for a band-limited f
f = plan.synthesis(x)
rt = plan.synthesis(plan.analysis(f))
e3 = float((rt - f).abs().pow(2).sum().sqrt() / f.abs().pow(2).sum().sqrt())
print(f"  dot-test synthesis pair {float(e1):.2e}, analysis pair {float(e2):.2e}, round trip {e3:.2e}")
assert e1 < 1e-10 and e2 < 1e-10 and e3 < 1e-10
print("OK")
