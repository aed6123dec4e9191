"""Import the UNMODIFIED reference package from /root/reference on top of the
oracle shims.  Only usable in the build container (the reference tree does not
exist on the GPU box); used by ``oracle/gen_golden.py`` to pin ``pxmcmc_ref``
and to generate ``tests/golden``.  TEST INFRASTRUCTURE ONLY."""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pxmcmc"))


def load():
    """Return the reference's ``pxmcmc`` as a namespace whose submodules are the
    reference's own files; pyssht/pys2let resolve to the oracle shims and
    healpy/astropy (setup-time only) to empty stubs."""
    if not available():
        raise RuntimeError("reference tree not present")
    from oracle.shims import pyssht as _pyssht, pys2let as _pys2let

    sys.modules.setdefault("pyssht", _pyssht)
    sys.modules.setdefault("pys2let", _pys2let)
    for name in ("healpy", "astropy", "astropy.coordinates"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            if name == "astropy.coordinates":
                mod.SkyCoord = None
            sys.modules[name] = mod
    if "pxmcmc" not in sys.modules or getattr(sys.modules["pxmcmc"], "__refloader__", False) is False:
        pkg = types.ModuleType("pxmcmc")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pxmcmc")]
        pkg.__refloader__ = True
        sys.modules["pxmcmc"] = pkg
    pkg = sys.modules["pxmcmc"]
    for sub in ("utils", "transforms", "measurements", "forward", "prior", "mcmc"):
        setattr(pkg, sub, importlib.import_module("pxmcmc." + sub))
    return pkg
