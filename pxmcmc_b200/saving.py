"""Saving an MCMC run, mirroring ``pxmcmc/saving.py`` of the reference: the tracked arrays become
datasets named ``logposterior, predictions, chain, L2s, priors, acceptances, deltas``, the runtime
parameters and any keyword arguments become attributes.

The reference writes HDF5 through h5py (pxmcmc/saving.py:18-36).  h5py is used when it can be
imported (same file, same names: the reference's plot scripts read it unchanged); otherwise the
same names go into ``<filename>.npz`` (datasets as arrays, attributes under ``attrs/<name>``) and
``load_mcmc`` reads either."""
import os

import numpy as np

DATASETS = (("logPi", "logposterior"), ("preds", "predictions"), ("chain", "chain"), ("L2s", "L2s"), ("priors", "priors"),
            ("acceptance_trace", "acceptances"), ("deltas_trace", "deltas"))


def _collect(mcmc, params, kwargs):
    data = {}
    for attr, name in DATASETS:
        if hasattr(mcmc, attr):
            arr = np.asarray(getattr(mcmc, attr))
            data[name] = arr.astype("i1") if name == "acceptances" else arr
    attrs = {k: getattr(params, k) for k in params.__dict__.keys()}
    attrs.update(kwargs)
    return data, attrs


def save_mcmc(mcmc, params, outpath, filename="outputs", **kwargs):
    """Save the tracked arrays of ``mcmc`` (after ``run``) and the parameters ``params`` in
    ``outpath/filename.hdf5`` (h5py available) or ``outpath/filename.npz``; returns the path."""
    data, attrs = _collect(mcmc, params, kwargs)
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        path = os.path.join(outpath, f"{filename}.hdf5")
        with h5py.File(path, "w") as f:
            for name, arr in data.items():
                f.create_dataset(name, data=arr)
            for k, v in attrs.items():
                f.attrs[k] = v
        return path
    path = os.path.join(outpath, f"{filename}.npz")
    flat = dict(data)
    for k, v in attrs.items():
        flat[f"attrs/{k}"] = np.asarray(v)
    np.savez(path, **flat)
    return path


def load_mcmc(path):
    """(datasets, attrs) dictionaries of a file written by ``save_mcmc``"""
    if path.endswith(".npz"):
        with np.load(path, allow_pickle=False) as f:
            data = {k: f[k] for k in f.files if not k.startswith("attrs/")}
            attrs = {k[6:]: (f[k][()] if f[k].ndim == 0 else f[k]) for k in f.files if k.startswith("attrs/")}
        return data, attrs
    import h5py

    with h5py.File(path, "r") as f:
        return {k: f[k][()] for k in f.keys()}, dict(f.attrs)
