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

  What is and is not pinned there (``tests/test_oracle_golden.py``):
  - the spin-weighted harmonics themselves: scipy ``sph_harm_y`` and closed-form Wigner d (independent of ssht);
  - the wavelet / scaling coefficients against their DEFINITION <f, R psi^j>, <f, R phi> by direct quadrature with
    the ``wavelet_tiling`` arrays the reference consumes
    (``test_wavelet_coefficients_are_the_defining_inner_products``): fixes the so3 (2 pi)^(-1/2), the tiling
    normalisations sqrt((2l+1)/8 pi^2) and sqrt((2l+1)/4 pi), grid positions and coefficient ordering;
  - exactly three numbers remain that only the wheels could confirm, each a single named symbol:
      1. ``s2let_ref.tiling_axisym: n = 300`` -- the trapezoid resolution of s2let's kappa integral (another
         quadrature changes kappa_j(l) at the 1e-6 level; partition of unity holds either way);
      2. ``s2let_ref._SQ2PI`` -- whether s2let stores W^j as defined (checked against the definition above) or
         rescaled by a constant (would rescale all wavelet coefficients, not the reconstruction);
      3. ``healpix_ref.map2alm(iter=3)`` -- healpy's default number of Jacobi refinements.
* ``greatcircle_ref`` -- restates ``greatcirclepaths==1.1.0`` (absent): points per radian, nearest-pixel binning and
  "average" weighting from the package's documented behaviour; parity unpinned, pinned by geometry and the
  path-average property the reference's test asserts.
"""
