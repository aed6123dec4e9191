"""CPU oracle for the pxmcmc hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the reported CPU baseline -- never as the thing that is shipped or measured as
the GPU path.  ``pxmcmc_b200`` never imports this package.

Parity status
-------------
* ``pxmcmc_ref``  -- restates the reference's own Python layer
  (``pxmcmc/mcmc.py``, ``forward.py``, ``prior.py``, ``measurements.py``,
  ``utils.py``).  PINNED: ``oracle/gen_golden.py`` runs the unmodified reference
  modules from ``/root/reference`` (on top of the shims below) and the
  committed fixtures in ``tests/golden/`` hold its outputs.
* ``ssht_ref`` / ``s2let_ref`` -- restate the published algorithms of the
  third-party C libraries the reference delegates to (``pyssht==1.5.2``,
  ``pys2let==2.2.6`` -- ``/root/reference/poetry.lock:1263-1265, 1234-1236``).
  Those wheels are absent from the image and cannot be installed (no network),
  and the reference's tests hold no golden numeric vector for them, so against
  the real binaries this part is **parity unpinned**.  It is pinned instead
  against every property the reference's tests assert for this path
  (``tests/test_transforms.py``, ``tests/test_measurements.py``,
  ``tests/test_utils.py:85-100``) and against closed-form spin-weighted
  spherical harmonics.
"""
