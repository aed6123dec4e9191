"""Development harness of the ring FFT: wavelet synthesis + its adjoint at the bench shape
(L=256, B=1.5, 64 chains) under every ring-FFT kernel choice; prints the ring-FFT device time per
pair of transforms (= one MYULA step's four launches) and the agreement with the multi-pass kernel."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pxmcmc_b200 import device as D
from pxmcmc_b200._lib import lib

L = int(os.environ.get("FFT_DEV_L", 256))
B = float(os.environ.get("FFT_DEV_B", 1.5))
NB = int(os.environ.get("FFT_DEV_NB", 64))
modes = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
plan = D.WaveletPlan.get(L, B, 2, NB)
g = torch.Generator(device="cuda").manual_seed(1)
coef = torch.randn((NB, plan.ncoefs), dtype=torch.float64, device="cuda", generator=g).to(torch.complex128)
coef = coef + 1j * torch.randn((NB, plan.ncoefs), dtype=torch.float64, device="cuda", generator=g)


def rel(a, b):
    return float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt())


ref = None
for mode in modes:
    lib.pxm_debug_set_fft_multipass(mode)
    for _ in range(3):
        pix = plan.synthesis(coef)
        back = plan.synthesis_adjoint(pix)
    torch.cuda.synchronize()
    reps = 10
    lib.pxm_profile_begin(4096)
    for _ in range(reps):
        pix = plan.synthesis(coef)
        back = plan.synthesis_adjoint(pix)
    ms = (C.c_double * 3)()
    cnt = (C.c_longlong * 3)()
    lib.pxm_profile_end(ms, cnt)
    line = f"mode {mode}: ring FFT {ms[1] / reps:.4f} ms per step ({cnt[1] // reps} launches), legendre {ms[0] / reps:.4f} ms"
    if ref is None:
        ref = (pix.clone(), back.clone())
    else:
        line += f"   rel-L2 vs mode {modes[0]}: pix {rel(pix, ref[0]):.2e}  coef {rel(back, ref[1]):.2e}"
    print(line, flush=True)
    clk = (C.c_ulonglong * 16)()
    if hasattr(lib, "pxm_debug_fft3_clocks") and lib.pxm_debug_fft3_clocks(clk) == 0 and mode == 3:
        tot = sum(clk)
        names = ["wait+bar0", "pass1", "bar1", "find+stage", "middle", "bar2", "pass3", "-"]
        for d in (0, 1):
            print(f"   fft3 DIR{d} phase cycles (thread 0 of each CTA, % of both): " + "  ".join(f"{n} {100 * c / tot:.1f}%" for n, c in zip(names, clk[8 * d:8 * d + 8]) if c), flush=True)
        nitems = 2 * NB * sum((l + 3) // 4 if l > 128 else (l + 7) // 8 for l in plan.bandlimits + [L] if l > 64) * reps
        print(f"   cycles per item (avg over {nitems} items): " + "  ".join(f"{n} {(clk[i] + clk[8 + i]) / nitems:.0f}" for i, n in enumerate(names) if clk[i]) + f"  total {tot / nitems:.0f}", flush=True)
lib.pxm_debug_set_fft_multipass(0)
