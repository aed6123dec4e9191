"""Golden vectors of the reference's uncertainty functions: runs the UNMODIFIED
/root/reference/pxmcmc/uncertainty.py (numpy quantiles; pyssht only supplies the MW grid sizes) on
seeded chains and writes tests/golden/ref_uncertainty.npz.  The chains themselves are regenerated
from the seed by the tests (see `make_chain`).  TEST INFRASTRUCTURE ONLY; build container only."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
CASES = ((200, 528, 11), (37, 130, 12), (5000, 96, 13))  # (nsamples, nparams, seed); 528 = ncoefs at L=10, B=2, J_min=2


def make_chain(nsamples, nparams, seed):
    """heavy-tailed samples rounded to 3 decimals so that ties occur"""
    rng = np.random.default_rng(seed)
    return np.round(rng.standard_t(3, size=(nsamples, nparams)) * np.linspace(0.1, 30.0, nparams), 3)


def main():
    from oracle import refloader

    refloader.load()
    unc = importlib.import_module("pxmcmc.uncertainty")
    d = {}
    for i, (n, p, seed) in enumerate(CASES):
        chain = make_chain(n, p, seed)
        d[f"ci_{i}"] = unc.credible_interval_range(chain)
        d[f"ci10_{i}"] = unc.credible_interval_range(chain, alpha=0.1)
    chain = make_chain(*CASES[0])
    maps = unc.wavelet_credible_interval_range(chain, 10, 2, 2)
    d["wav_shapes"] = np.array([m.shape for m in maps])
    d["wav_flat"] = np.concatenate([m.ravel() for m in maps])
    logpis = -np.abs(make_chain(1, 400, 14)[0])
    d["threshold"] = unc.credible_region_threshold(logpis)
    d["threshold20"] = unc.credible_region_threshold(logpis, alpha=0.2)
    np.savez(os.path.join(OUT, "ref_uncertainty.npz"), **d)
    print("written", os.path.join(OUT, "ref_uncertainty.npz"), os.path.getsize(os.path.join(OUT, "ref_uncertainty.npz")), "B")


if __name__ == "__main__":
    main()
