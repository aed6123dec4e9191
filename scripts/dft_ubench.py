import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pxmcmc_b200 import _lib
_lib.ensure_device()
lib = _lib.lib
iters = 2000
for ctas in (1, 2):
    ms = C.c_float()
    lib.pxm_debug_dft_ubench(ctas, iters, C.byref(ms))
    cyc = ms.value * 1e-3 * 1.965e9
    # FP64 instructions per iteration per thread: 2 x 460 (dft32) + 128 (cmul)
    n64 = 2 * 460 + 128
    print(f"{ctas} CTA/SM (= {ctas} warp per scheduler): {ms.value:.3f} ms, {cyc / iters:.0f} cycles per iteration, "
          f"{cyc / iters / n64:.2f} cycles per FP64 instruction per warp, pipe busy ~{100 * 2 * n64 * ctas / (cyc / iters):.0f} %")
