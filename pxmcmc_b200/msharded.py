"""m-sharded transforms: ONE chain at a bandlimit too large or too slow for one
GPU, spread over the GPUs of an NVLink/NVSwitch node (the
reference itself is single-process, there is nothing to mirror).

Partition (identical on every rank, pure functions of (L, world)):

* harmonic space -- rank r owns the azimuthal orders |m| with
  ``owner_of_m(|m|, world) == r`` (snake order: the triangular work L-|m|
  balances).  ``flm`` arrays keep their full length, a rank reads/writes only its
  own orders (the others read as 0 on output).  Legendre tables: 1/world per GPU.
* pixel / coefficient space -- of every ring grid (pixel map, each wavelet scale
  map) a rank owns the rows ``ring_range(...)``, whole blocks of 64 rings.  A
  *local vector* is the concatenation, in the reference's scale order, of the
  owned rows of each map; all elementwise kernels (prox, Langevin update,
  residual, masks) run unchanged on local vectors.

The theta<->m transposition is not a separate collective: the Legendre
contraction kernel stores its output tiles directly into the owner's ring buffer
(peer memory over NVLink) and pulls ring blocks from their owners with the same
``cp.async.bulk`` copies it uses locally; a flag-based peer barrier kernel
separates local phases from peer-access phases.  torch.distributed is used only
to exchange the CUDA-IPC handles at set-up and for scalar reductions.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import device as D
from ._lib import check, lib, ptr


# ----------------------------------------------------------------------------
# the partition (host-only; also what the CPU tests check)
# ----------------------------------------------------------------------------
def owner_of_m(abs_m, world):
    return lib.pxm_shard_owner_of_m(int(abs_m), int(world))


def ring_range(ell, rot, rank, world):
    t0, t1 = C.c_int(), C.c_int()
    check(lib.pxm_shard_ring_range(int(ell), int(rot), int(rank), int(world), C.byref(t0), C.byref(t1)))
    return t0.value, t1.value


def flm_owner_mask(L, rank, world):
    """bool[L*L]: the harmonic coefficients (index l*l+l+m) this rank owns"""
    mask = np.zeros(L * L, dtype=bool)
    for el in range(L):
        for m in range(-el, el + 1):
            mask[el * el + el + m] = owner_of_m(abs(m), world) == rank
    return mask


class MapLayout:
    """Rows of a sequence of MW maps (bandlimits ``ells``, ownership rotations
    ``rots``) owned by ``rank``; converts between full and local vectors."""

    def __init__(self, ells, rots, rank, world):
        self.ells, self.rots, self.rank, self.world = list(ells), list(rots), rank, world
        self.rows = [ring_range(e, r, rank, world) for e, r in zip(self.ells, self.rots)]
        self.full_off = np.concatenate([[0], np.cumsum([e * (2 * e - 1) for e in self.ells])]).astype(np.int64)
        self.n_full = int(self.full_off[-1])
        idx = []
        for (t0, t1), e, off in zip(self.rows, self.ells, self.full_off[:-1]):
            n = 2 * e - 1
            idx.append(off + np.arange(t0 * n, t1 * n, dtype=np.int64))
        self.index = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)  # local -> full position
        self.n_local = int(self.index.size)

    def to_local(self, full):
        full = np.asarray(full)
        return np.ascontiguousarray(full[..., self.index])

    def scatter_into(self, full, local):
        full[..., self.index] = np.asarray(local)
        return full


# ----------------------------------------------------------------------------
# rank groups: how workspace addresses travel
# ----------------------------------------------------------------------------
class ProcessGroupExchange:
    """one process per GPU (torch.distributed initialised): CUDA IPC handles via all_gather_object"""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._opened = []

    def exchange(self, ws_ptr):
        h = (C.c_ubyte * 64)()
        check(lib.pxm_ipc_export(C.c_void_p(ws_ptr), h))
        handles = [None] * self.world
        self.dist.all_gather_object(handles, bytes(h), group=self.group)
        out = (C.c_void_p * self.world)()
        for q, hb in enumerate(handles):
            if q == self.rank:
                out[q] = ws_ptr
            else:
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(hb)
                check(lib.pxm_ipc_open(buf, C.byref(p)))
                self._opened.append(p)
                out[q] = p.value
        return out

    def ready(self):
        """every rank has attached: nobody pushes into a workspace that is not mapped yet"""
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)


class _ShardedBase:
    kind = None

    def _attach(self, exchange):
        # tables now, not on first use: their generator synchronises the device (cudaFree),
        # which must not happen while a peer is waiting in the flag barrier
        check(getattr(lib, f"pxm_{self.kind}_plan_prepare")(self.h))
        nbytes = C.c_size_t()
        ws = getattr(lib, f"pxm_{self.kind}_plan_workspace")(self.h, C.byref(nbytes))
        self.ws_ptr, self.ws_bytes = ws, nbytes.value
        if exchange is not None:
            peers = exchange.exchange(ws)
            check(getattr(lib, f"pxm_{self.kind}_plan_attach")(self.h, peers))
            exchange.ready()

    def attach_pointers(self, ptrs):
        arr = (C.c_void_p * self.world)(*ptrs)
        check(getattr(lib, f"pxm_{self.kind}_plan_attach")(self.h, arr))

    def barrier_ok(self):
        e = C.c_longlong()
        check(getattr(lib, f"pxm_{self.kind}_plan_barrier_status")(self.h, C.byref(e)))
        return e.value == 0

    def _call(self, fn, x, n_in, n_out, extra=()):
        x2, was1 = D.batch2d(x)
        if x2.shape[1] != n_in:
            raise ValueError(f"expected local vectors of length {n_in}, got {x2.shape[1]}")
        out = torch.empty((x2.shape[0], n_out), dtype=D.CDT, device=x2.device)
        check(fn(self.h, ptr(x2), ptr(out), x2.shape[0], *extra, _lib.stream_ptr()))
        return out[0] if was1 else out


class ShardedWaveletPlan(_ShardedBase):
    """pxm_wav_plan for rank `rank` of `world`; vectors are LOCAL (see module docstring)."""

    kind = "wav"

    def __init__(self, L, B, J_min, rank, world, nbatch=1, exchange=None):
        D.dev()
        self.L, self.B, self.J_min, self.rank, self.world = int(L), float(B), int(J_min), int(rank), int(world)
        h = C.c_void_p()
        check(lib.pxm_wav_plan_create_sharded(self.L, self.B, self.J_min, int(nbatch), self.rank, self.world, C.byref(h)))
        self.h = h
        ns, nc, nsc, J, tb = C.c_int(), C.c_longlong(), C.c_longlong(), C.c_int(), C.c_longlong()
        check(lib.pxm_wav_plan_info(h, C.byref(ns), C.byref(nc), C.byref(nsc), C.byref(J), C.byref(tb)))
        self.nscales_total, self.ncoefs, self.table_bytes = ns.value, nc.value, tb.value
        bl = (C.c_int * ns.value)()
        check(lib.pxm_wav_plan_bandlimits(h, bl, ns.value))
        self.bandlimits = list(bl)
        t0 = (C.c_int * (ns.value + 1))()
        t1 = (C.c_int * (ns.value + 1))()
        ncl, npl = C.c_longlong(), C.c_longlong()
        check(lib.pxm_wav_plan_local_rows(h, t0, t1, C.byref(ncl), C.byref(npl)))
        self.ncoefs_local, self.npix_local = ncl.value, npl.value
        self.coef_layout = MapLayout(self.bandlimits, range(1, ns.value + 1), self.rank, self.world)
        self.pix_layout = MapLayout([self.L], [0], self.rank, self.world)
        assert self.coef_layout.rows == [(t0[i], t1[i]) for i in range(ns.value)]
        assert self.pix_layout.rows == [(t0[ns.value], t1[ns.value])]
        assert self.coef_layout.n_local == self.ncoefs_local and self.pix_layout.n_local == self.npix_local
        self._attach(exchange)

    def synthesis(self, coef_local):
        return self._call(lib.pxm_wav_synthesis, coef_local, self.ncoefs_local, self.npix_local)

    def synthesis_adjoint(self, pix_local):
        return self._call(lib.pxm_wav_synthesis_adjoint, pix_local, self.npix_local, self.ncoefs_local)

    def analysis(self, pix_local):
        return self._call(lib.pxm_wav_analysis, pix_local, self.npix_local, self.ncoefs_local)

    def analysis_adjoint(self, coef_local):
        return self._call(lib.pxm_wav_analysis_adjoint, coef_local, self.ncoefs_local, self.npix_local)

    def synthesis_harmonic(self, coef_local):
        """local coefficients -> f_lm (full length, only this rank's orders populated)"""
        return self._call(lib.pxm_wav_synthesis_harmonic, coef_local, self.ncoefs_local, self.L * self.L)

    def synthesis_adjoint_harmonic(self, flm):
        return self._call(lib.pxm_wav_synthesis_adjoint_harmonic, flm, self.L * self.L, self.ncoefs_local)


class ShardedShtPlan(_ShardedBase):
    """pxm_sht_plan for rank `rank` of `world`: pixel vectors are the local rows, flm
    vectors have full length with only the owned orders populated."""

    kind = "sht"

    def __init__(self, L, spin, rank, world, nbatch=1, exchange=None):
        D.dev()
        self.L, self.spin, self.rank, self.world = int(L), int(spin), int(rank), int(world)
        h = C.c_void_p()
        check(lib.pxm_sht_plan_create_sharded(self.L, self.spin, int(nbatch), self.rank, self.world, C.byref(h)))
        self.h = h
        t0, t1 = C.c_int(), C.c_int()
        check(lib.pxm_sht_plan_local_rows(h, C.byref(t0), C.byref(t1)))
        self.pix_layout = MapLayout([self.L], [0], self.rank, self.world)
        assert self.pix_layout.rows == [(t0.value, t1.value)]
        self.npix_local = self.pix_layout.n_local
        self.nlm = self.L * self.L
        self.table_bytes = int(lib.pxm_sht_plan_table_bytes(h))  # Lambda + W families, this rank's orders
        self._attach(exchange)

    def inverse(self, flm, gl=None):
        return self._call(lib.pxm_sht_inverse, flm, self.nlm, self.npix_local, (ptr(gl),))

    def forward(self, f_local, gl=None):
        return self._call(lib.pxm_sht_forward, f_local, self.npix_local, self.nlm, (ptr(gl),))

    def inverse_adjoint(self, f_local, gl=None):
        return self._call(lib.pxm_sht_inverse_adjoint, f_local, self.npix_local, self.nlm, (ptr(gl),))

    def forward_adjoint(self, flm, gl=None):
        return self._call(lib.pxm_sht_forward_adjoint, flm, self.nlm, self.npix_local, (ptr(gl),))


# ----------------------------------------------------------------------------
# all ranks inside ONE process on ONE GPU (one CUDA stream per rank): exercises the
# very same sharded code path -- peer pointers, push/pull contractions, flag
# barrier -- where only a single GPU is available (the driver's `pytest -m gpu`).
# ----------------------------------------------------------------------------
class SimulatedRanks:
    def __init__(self, world, *factories):
        """factories: make_plan(rank, world) -> a sharded plan created with exchange=None.
        `plan_sets[k][r]` is rank r's plan from factory k; `plans` = `plan_sets[0]`."""
        self.world = world
        self.plan_sets = []
        for make_plan in factories:
            plans = [make_plan(r, world) for r in range(world)]
            ptrs = [p.ws_ptr for p in plans]
            for p in plans:
                p.attach_pointers(ptrs)
            self.plan_sets.append(plans)
        self.plans = self.plan_sets[0]
        self.streams = [torch.cuda.Stream() for _ in range(world)]
        torch.cuda.synchronize()

    def map(self, fn):
        """fn(rank) is called for every rank with that rank's stream current; it must only
        enqueue work (no host synchronisation: the other ranks' kernels are not queued yet).
        Returns the per-rank results after every stream has drained."""
        outs = []
        cur = torch.cuda.current_stream()
        for r, st in enumerate(self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(fn(r))
        for st in self.streams:
            cur.wait_stream(st)
        torch.cuda.synchronize()
        for plans in self.plan_sets:
            for p in plans:
                if not p.barrier_ok():
                    raise RuntimeError("peer barrier timed out")
        return outs

    def run(self, method, inputs, **kw):
        """plan.<method>(inputs[r]) for every rank of the first plan set"""
        return self.map(lambda r: getattr(self.plans[r], method)(inputs[r], **kw))


# ----------------------------------------------------------------------------
# sharded counterparts of the reference-facing classes: same method names, LOCAL
# vectors.  They plug into pxmcmc_b200.forward.ForwardOperator and the samplers of
# pxmcmc_b200.mcmc unchanged (which only see vectors of the local length).
# ----------------------------------------------------------------------------
class ShardedSphericalWaveletTransform:
    """`SphericalWaveletTransform` (pxmcmc/transforms.py:59-166) of one rank: `inverse`,
    `inverse_adjoint`, `forward`, `forward_adjoint` on local coefficient / pixel vectors."""

    _pxm_native = True

    def __init__(self, L, B, J_min, rank, world, exchange=None, nchains=1, plan=None):
        self.plan = plan if plan is not None else ShardedWaveletPlan(L, B, J_min, rank, world, nbatch=nchains, exchange=exchange)
        self.L, self.B, self.J_min = L, B, J_min
        self.bandlimits = list(self.plan.bandlimits)
        self.ncoefs = self.plan.ncoefs_local          # local sizes: what the sampler sees
        self.ncoefs_global = self.plan.ncoefs
        self.coef_layout, self.pix_layout = self.plan.coef_layout, self.plan.pix_layout

    def _apply(self, name, X):
        return D.like_input(getattr(self.plan, name)(D.to_dev_c(X)), X)

    def forward(self, X):
        return self._apply("analysis", X)

    def inverse(self, X):
        return self._apply("synthesis", X)

    def inverse_adjoint(self, X):
        return self._apply("synthesis_adjoint", X)

    def forward_adjoint(self, X):
        return self._apply("analysis_adjoint", X)

    spin = 0

    def _inverse_harmonic(self, X):
        return self._apply("synthesis_harmonic", X)

    def _inverse_adjoint_harmonic(self, flm):
        return self._apply("synthesis_adjoint_harmonic", flm)


class ShardedWeakLensing:
    """`WeakLensing` (pxmcmc/measurements.py:185-304) of one rank: convergence rows in,
    masked + covariance-weighted shear of the same rows out, and the adjoint chain.
    The harmonic coefficients in between stay m-sharded; the kappa->gamma kernel is the
    per-degree multiplier of the spin-2 transform."""

    _pxm_native = True

    def __init__(self, L, rank, world, mask=None, ngal=None, exchange=None, plans=None):
        from .measurements import WeakLensing

        host = WeakLensing(L, mask=mask, ngal=ngal)  # host-side set-up only (mask, inv_cov, kernel)
        self.L, self.rank, self.world = L, rank, world
        self.s0, self.s2 = plans if plans is not None else (
            ShardedShtPlan(L, 0, rank, world, exchange=exchange), ShardedShtPlan(L, 2, rank, world, exchange=exchange))
        self.pix_layout = self.s0.pix_layout
        (t0, t1), n = self.pix_layout.rows[0], 2 * L - 1
        local_mask = host.mask[t0:t1].ravel()
        self.npix = self.pix_layout.n_local
        self.ndata = int(local_mask.sum())
        # position of the local data inside the reference's full data vector (C-order boolean gather)
        before = int(host.mask[:t0].sum())
        self.data_index = before + np.arange(self.ndata, dtype=np.int64)
        self.ndata_global = int(host.mask.sum())
        self.inv_cov = np.asarray(host.inv_cov)[self.data_index]
        dv = D.dev()
        self._idx = torch.from_numpy(np.flatnonzero(local_mask).astype(np.int32)).to(dv)
        self._w = D.to_dev_f(self.inv_cov)
        self._gl = D.to_dev_f(host._kernel_per_l())

    def forward(self, kappa):
        x = D.to_dev_c(kappa)
        klm = self.s0.forward(x)
        gamma = self.s2.inverse(klm, gl=self._gl)
        return D.like_input(D.gather_dev(gamma, self._idx, self._w, self.ndata), kappa)

    def adjoint(self, gamma):
        y = D.to_dev_c(gamma)
        g = D.scatter_dev(y, self._idx, self._w, self.npix)
        glm = self.s2.inverse_adjoint(g, gl=self._gl)
        return D.like_input(self.s0.forward_adjoint(glm), gamma)

    # harmonic-space input / output (see measurements.WeakLensing): the f_lm stay m-sharded between the
    # wavelet plan and the spin-2 plan, which deal the orders to the ranks with the same owner_of_m
    _pxm_harmonic_input = True

    def _forward_from_harmonic(self, klm):
        gamma = self.s2.inverse(D.to_dev_c(klm), gl=self._gl)
        return D.like_input(D.gather_dev(gamma, self._idx, self._w, self.ndata), klm)

    def _adjoint_to_harmonic(self, gamma):
        g = D.scatter_dev(D.to_dev_c(gamma), self._idx, self._w, self.npix)
        return D.like_input(self.s2.inverse_adjoint(g, gl=self._gl), gamma)


def sharded_s2_wavelets_l1(transform, T, L, B, J_min, cls=None, **kw):
    """The reference's `S2_Wavelets_L1` (or a subclass) restricted to this rank's
    coefficients: thresholds and prior weights are the full vectors gathered at the
    local positions, so prox and Langevin update are element-for-element those of the
    unsharded sampler."""
    from . import prior as P

    full = (cls or P.S2_Wavelets_L1)("synthesis", transform.inverse, transform.inverse_adjoint, T, L, B, J_min, **kw)
    idx = transform.coef_layout.index
    full.T = np.ascontiguousarray(np.asarray(full.T)[idx])
    full.map_weights = np.ascontiguousarray(np.asarray(full.map_weights)[idx])
    full._Tdev = full._wdev = None
    return full


def allreduce_sum(group=None):
    """scalar reductions across ranks (log posterior, L2, prior): what the samplers call
    through `forward._pxm_allreduce` when the operator is sharded"""
    import torch.distributed as dist

    def f(t):
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t

    return f


# the harmonic-space composition in ForwardOperator bypasses `forward` / `adjoint`: it is only taken while a
# (sub)class still uses these very implementations
ShardedWeakLensing._pxm_fused_methods = (ShardedWeakLensing.forward, ShardedWeakLensing.adjoint)
