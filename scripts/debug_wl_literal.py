"""localise the disagreement between the literal and the harmonic-space composition of config 4 (L=512)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from pxmcmc_b200 import device as D, sht, measurements, transforms
from test_gpu_configs import wl_mask

def rl(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

for L in [int(v) for v in (sys.argv[1:] or ["257", "384", "512"])]:
    B, J = 2.0, 2
    mask = wl_mask(L)
    wl = measurements.WeakLensing(L, mask=mask, ngal=np.full((L, 2 * L - 1), 30.0))
    tr = transforms.SphericalWaveletTransform(L, B, J)
    rng = np.random.default_rng(1)
    y = rng.standard_normal(wl.ndata) + 1j * rng.standard_normal(wl.ndata)
    yd = D.to_dev_c(y)
    glm = wl._adjoint_to_harmonic(yd)                      # scatter + spin-2 inverse_adjoint
    g_fused = tr._inverse_adjoint_harmonic(glm)
    kap = wl.adjoint(yd)                                   # ... + spin-0 forward_adjoint
    g_lit = tr.inverse_adjoint(kap)
    print(f"L={L}: gradg literal vs fused {rl(g_lit.cpu().numpy(), g_fused.cpu().numpy()):.2e}", flush=True)
    # pieces
    kap2 = D.ShtPlan.get(L, 0, 1).forward_adjoint(glm)
    print("   wl.adjoint vs explicit forward_adjoint(glm):", rl(kap.cpu().numpy(), kap2.cpu().numpy()))
    back = D.ShtPlan.get(L, 0, 1).inverse_adjoint(kap2)    # A_inv^dagger A_fwd^dagger = I
    print("   inverse_adjoint(forward_adjoint(glm)) vs glm:", rl(back.cpu().numpy(), glm.cpu().numpy()))
    g_lit2 = tr.inverse_adjoint(D.to_dev_c(kap2.cpu().numpy()))
    print("   literal via a host round trip vs literal:", rl(g_lit2.cpu().numpy(), g_lit.cpu().numpy()))
    # per scale
    a, b = g_lit.cpu().numpy().ravel(), g_fused.cpu().numpy().ravel()
    off = 0
    for bl in tr.bandlimits:
        n = bl * (2 * bl - 1)
        print(f"     scale bandlimit {bl}: {rl(a[off:off+n], b[off:off+n]):.2e}  |fused| {np.linalg.norm(b[off:off+n]):.3e}")
        off += n
    # norms
    print("   |kappa| =", np.linalg.norm(kap.cpu().numpy()), " |glm| =", np.linalg.norm(glm.cpu().numpy()))
    # forward direction
    x = rng.standard_normal(tr.ncoefs) + 1j * rng.standard_normal(tr.ncoefs)
    xd = D.to_dev_c(x)
    p_fused = wl._forward_from_harmonic(tr._inverse_harmonic(xd))
    p_lit = wl.forward(tr.inverse(xd))
    print(f"   forward literal vs fused {rl(p_lit.cpu().numpy(), p_fused.cpu().numpy()):.2e}", flush=True)
    # random pixel map through synthesis_adjoint twice (determinism)
    f = rng.standard_normal(L * (2 * L - 1)) + 1j * rng.standard_normal(L * (2 * L - 1))
    fd = D.to_dev_c(f)
    r1 = tr.inverse_adjoint(fd).cpu().numpy(); r2 = tr.inverse_adjoint(fd).cpu().numpy()
    print("   determinism of synthesis_adjoint:", rl(r1, r2))
    # kappa-like input: is it the size of the numbers?
    big = tr.inverse_adjoint(D.to_dev_c(kap2.cpu().numpy() * 1e-6)).cpu().numpy() * 1e6
    print("   scaled input (1e-6) literal vs literal:", rl(big, g_lit.cpu().numpy()))
