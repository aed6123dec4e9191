"""Data ingestion for the reference's experiment drivers without healpy / astropy:

* ``read_map`` / ``write_map`` -- HEALPix maps in FITS binary tables (``hp.read_map``,
  experiments/earthtopography/main.py:80, experiments/weaklensing/main.py:31);
* ``smoothing`` -- ``hp.smoothing`` (experiments/weaklensing/main.py:36): Gaussian beam applied to the a_lm through
  the device HEALPix transforms (``pxm_hpx_*``);
* ``load_gammas`` -- the weak-lensing driver's data preparation (experiments/weaklensing/main.py:23-39);
* ``build_mask`` -- galactic-plane + ecliptic mask on the MW grid (pxmcmc/utils.py:320-349) with the fixed
  ICRS -> galactic rotation instead of astropy's ``SkyCoord``.

All of this is set-up work that runs once per experiment; the transforms it needs are the library's kernels.
"""
import numpy as np

from . import utils

_BLOCK = 2880
_TFORM = {"E": ">f4", "D": ">f8", "J": ">i4", "I": ">i2", "K": ">i8", "B": "u1"}


# ---------------------------------------------------------------------------------------------------------
# FITS binary tables holding HEALPix maps
# ---------------------------------------------------------------------------------------------------------
def _read_header(buf, pos):
    cards = {}
    while True:
        block = buf[pos:pos + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        pos += _BLOCK
        done = False
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                v = val.lstrip()[1:]
                v = v[: v.index("'")] if "'" in v else v
                cards[key] = v.strip()
            else:
                v = val.split("/")[0].strip()
                if v in ("T", "F"):
                    cards[key] = v == "T"
                else:
                    try:
                        cards[key] = int(v)
                    except ValueError:
                        try:
                            cards[key] = float(v.replace("D", "E"))
                        except ValueError:
                            cards[key] = v
        if done:
            return cards, pos


def nest2ring_order(nside):
    """index array `r` with ring_map[r[p]] = nested_map[p] (HEALPix primer, section 4.1: the NESTED index interleaves
    the bits of (x, y) inside one of 12 base faces)"""
    npix = 12 * nside * nside
    p = np.arange(npix, dtype=np.int64)
    face, pf = p // (nside * nside), p % (nside * nside)
    ix = np.zeros(npix, dtype=np.int64)
    iy = np.zeros(npix, dtype=np.int64)
    for b in range(int(np.log2(nside)) if nside > 1 else 0):
        ix |= ((pf >> (2 * b)) & 1) << b
        iy |= ((pf >> (2 * b + 1)) & 1) << b
    jrll = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
    jpll = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])
    jr = jrll[face] * nside - ix - iy - 1  # ring number in 1 .. 4 nside - 1
    nr = np.where(jr < nside, jr, np.where(jr > 3 * nside, 4 * nside - jr, nside))
    n_before = np.where(jr < nside, 2 * jr * (jr - 1),
                        np.where(jr > 3 * nside, npix - 2 * (4 * nside - jr + 1) * (4 * nside - jr),
                                 2 * nside * (nside - 1) + (jr - nside) * 4 * nside))
    kshift = np.where((jr >= nside) & (jr <= 3 * nside), (jr - nside) & 1, 0)
    jp = (jpll[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > 4 * nr, jp - 4 * nr, jp)
    jp = np.where(jp < 1, jp + 4 * nr, jp)
    return n_before + jp - 1


def read_map(filename, field=0, nest=False, dtype=np.float64, h=False, **kwargs):
    """``healpy.read_map``: column `field` of the first binary-table extension, returned in RING order (or NESTED when
    `nest=True`), whatever the file's ORDERING.  Implicit full-sky maps only (what the reference's inputs are)."""
    with open(filename, "rb") as f:
        buf = f.read()
    hdr, pos = _read_header(buf, 0)
    if not hdr.get("SIMPLE", False):
        raise ValueError("not a FITS file")
    pos += ((abs(hdr.get("BITPIX", 8)) // 8 * int(np.prod([hdr.get(f"NAXIS{i + 1}", 0) for i in range(hdr.get("NAXIS", 0))]))
             if hdr.get("NAXIS", 0) else 0) + _BLOCK - 1) // _BLOCK * _BLOCK
    while True:
        ext, pos = _read_header(buf, pos)
        nbytes = ext.get("NAXIS1", 0) * ext.get("NAXIS2", 0) + ext.get("PCOUNT", 0)
        if ext.get("XTENSION", "").startswith("BINTABLE"):
            break
        pos += (nbytes + _BLOCK - 1) // _BLOCK * _BLOCK
    if ext.get("INDXSCHM", "IMPLICIT").strip() != "IMPLICIT" or ext.get("OBJECT", "FULLSKY").strip() == "PARTIAL":
        raise NotImplementedError("explicitly indexed / partial-sky HEALPix files are not supported")
    fields, offset = [], 0
    for k in range(1, ext["TFIELDS"] + 1):
        tf = ext[f"TFORM{k}"].strip()
        code = tf[-1]
        rep = int(tf[:-1]) if tf[:-1] else 1
        if code not in _TFORM:
            raise NotImplementedError(f"FITS column format {tf}")
        dt = np.dtype(_TFORM[code])
        fields.append((offset, rep, dt))
        offset += rep * dt.itemsize
    if offset != ext["NAXIS1"]:
        raise ValueError("FITS table row length does not match its column formats")
    rows = np.frombuffer(buf, dtype=np.uint8, count=ext["NAXIS1"] * ext["NAXIS2"], offset=pos).reshape(ext["NAXIS2"], ext["NAXIS1"])
    sel = [field] if np.ndim(field) == 0 else list(field)
    maps = []
    for fi in sel:
        o, rep, dt = fields[fi]
        col = np.ascontiguousarray(rows[:, o:o + rep * dt.itemsize]).view(dt).reshape(-1).astype(dtype)
        nside = ext.get("NSIDE") or utils._nside_of(col.size)
        if col.size != 12 * nside * nside:
            raise ValueError("column length is not 12 nside^2")
        in_nest = ext.get("ORDERING", "RING").strip().upper().startswith("NEST")
        if in_nest != bool(nest):
            order = nest2ring_order(nside)
            if in_nest:
                ring = np.empty_like(col)
                ring[order] = col
                col = ring
            else:
                col = col[order]
        maps.append(col)
    out = maps[0] if np.ndim(field) == 0 else np.array(maps)
    return (out, sorted(ext.items())) if h else out


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = ("T" if value else "F").rjust(20)
    elif isinstance(value, (int, np.integer)):
        v = str(int(value)).rjust(20)
    elif isinstance(value, float):
        v = repr(value).rjust(20)
    else:
        v = ("'" + str(value).ljust(8) + "'").ljust(20)
    return f"{key:<8}= {v}{(' / ' + comment) if comment else ''}".ljust(80)[:80]


def write_map(filename, m, nest=False, dtype=np.float32, column_name="T", overwrite=True, **kwargs):
    """``healpy.write_map`` layout (the one of ETOPO1_Ice_hpx_256.fits): an empty primary HDU and one BINTABLE whose
    single column holds 1024 pixels per row."""
    m = np.asarray(m)
    nside = utils._nside_of(m.size)
    rep = 1024 if m.size % 1024 == 0 else 1
    code = {np.dtype(np.float32): "E", np.dtype(np.float64): "D"}[np.dtype(dtype)]
    be = np.dtype(_TFORM[code])
    data = m.astype(be).tobytes()
    primary = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True), "END".ljust(80)]
    ext = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8), _card("NAXIS", 2),
           _card("NAXIS1", rep * be.itemsize), _card("NAXIS2", m.size // rep), _card("PCOUNT", 0), _card("GCOUNT", 1),
           _card("TFIELDS", 1), _card("TTYPE1", column_name), _card("TFORM1", f"{rep}{code}"), _card("PIXTYPE", "HEALPIX"),
           _card("ORDERING", "NESTED" if nest else "RING"), _card("EXTNAME", "xtension"), _card("NSIDE", nside),
           _card("FIRSTPIX", 0), _card("LASTPIX", m.size - 1), _card("INDXSCHM", "IMPLICIT"), _card("OBJECT", "FULLSKY"),
           "END".ljust(80)]

    def pad(b, fill):
        return b + fill * ((-len(b)) % _BLOCK)

    with open(filename, "wb" if overwrite else "xb") as f:
        f.write(pad("".join(primary).encode("ascii"), b" "))
        f.write(pad("".join(ext).encode("ascii"), b" "))
        f.write(pad(data, b"\0"))


# ---------------------------------------------------------------------------------------------------------
# hp.smoothing and the weak-lensing data preparation
# ---------------------------------------------------------------------------------------------------------
def gauss_beam(sigma, lmax):
    """healpy.gauss_beam (temperature): b_l = exp(-l (l + 1) sigma^2 / 2), sigma in radians"""
    el = np.arange(lmax + 1, dtype=float)
    return np.exp(-0.5 * el * (el + 1) * sigma * sigma)


def almxfl(alm, fl, lmax):
    """healpy.almxfl on healpy's m-major storage of the m >= 0 coefficients"""
    alm = np.array(alm, dtype=complex)
    for m in range(lmax + 1):
        els = np.arange(m, lmax + 1)
        alm[utils.alm_hp_index(els, m, lmax)] *= fl[els]
    return alm


def smoothing(map_in, fwhm=0.0, sigma=None, lmax=None, iter=3, **kwargs):
    """``healpy.smoothing`` of a RING-ordered real map: map2alm (``iter`` Jacobi refinements) -> Gaussian beam -> alm2map.
    healpy's default bandlimit is 3 nside - 1; the dense Legendre tables of the device transform make that the default
    only up to lmax = 1023 -- pass `lmax` for finer maps (a map that is already bandlimited, as in the reference's
    driver, loses nothing)."""
    map_in = np.asarray(map_in, dtype=float)
    nside = utils._nside_of(map_in.size)
    if sigma is None:
        sigma = fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    if lmax is None:
        lmax = 3 * nside - 1
        if lmax > 1023:
            raise NotImplementedError("smoothing: pass lmax (<= 1023) for nside > 341")
    alm = almxfl(utils.map2alm(map_in, lmax, iter=iter), gauss_beam(sigma, lmax), lmax)
    return utils.alm2map(alm, nside)


def load_gammas(kappa_fits_file, L, wl, sigma=np.radians(50 / 60)):
    """Shear data of the weak-lensing experiment from a HEALPix convergence map
    (experiments/weaklensing/main.py:23-39): bandlimit the map at lmax = L - 1, smooth it with a 50-arcmin Gaussian,
    resample it on the MW grid, apply the masked Kaiser-Squires operator `wl`.

    The reference goes through an intermediate HEALPix map at nside = 3 lmax - 1 between the three steps
    (alm2map -> smoothing -> map2alm); the signal is bandlimited at lmax throughout, so the beam is applied to the
    a_lm directly here (the intermediate maps only add healpy's quadrature error)."""
    kappa = read_map(kappa_fits_file)
    lmax = L - 1
    alm = almxfl(utils.map2alm(kappa, lmax), gauss_beam(sigma, lmax), lmax)
    kappa_mw = utils.alm2map_mw(utils.lm_hp2lm(alm, L), L)
    return wl.forward(kappa_mw)


# ---------------------------------------------------------------------------------------------------------
# Euclid-like mask
# ---------------------------------------------------------------------------------------------------------
def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    m = {0: [[1, 0, 0], [0, c, s], [0, -s, c]], 1: [[c, 0, -s], [0, 1, 0], [s, 0, c]], 2: [[c, s, 0], [-s, c, 0], [0, 0, 1]]}[axis]
    return np.array(m, dtype=float)


def icrs_to_galactic_matrix():
    """The constant rotation astropy applies for ``SkyCoord(...).transform_to("galactic")`` from ICRS: the ICRS -> FK5
    (J2000) frame bias (eta0 = -19.9 mas, xi0 = 9.1 mas, da0 = -22.9 mas; IERS Conventions 2003) followed by the IAU
    1958 definition of the galactic system referred to J2000 (north galactic pole at RA 192.85948 deg, Dec 27.12825 deg,
    position angle of the galactic centre 122.93192 deg)."""
    mas = np.radians(1.0 / 3600e3)
    bias = _rot(0, -(-19.9) * mas) @ _rot(1, 9.1 * mas) @ _rot(2, -22.9 * mas)
    ngp_ra, ngp_dec, lon0 = np.radians(192.85948), np.radians(27.12825), np.radians(122.93192)
    fk5_to_gal = _rot(2, np.pi - lon0) @ _rot(1, np.pi / 2 - ngp_dec) @ _rot(2, ngp_ra)
    return fk5_to_gal @ bias


def build_mask(L, size=20):
    """Mask of the galactic plane and the ecliptic on the MW grid, 0 where masked (pxmcmc/utils.py:320-349): rings
    with |90 - theta| < size degrees, and pixels whose galactic latitude is below `size` degrees when the MW grid is
    read as ICRS coordinates (ra, dec) = (phi - 180, theta - 90) degrees -- the reference's own convention."""
    thetas, phis = utils.mw_sample_positions(L)
    mask = np.ones((L, 2 * L - 1))
    mask[np.abs(90 - np.degrees(thetas)) < size, :] = 0
    dec = np.radians(np.degrees(thetas) - 90)[:, None]
    ra = np.radians(np.degrees(phis) - 180)[None, :]
    v = np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec) * np.ones_like(ra)])
    g = np.tensordot(icrs_to_galactic_matrix(), v, axes=1)
    b = np.degrees(np.arcsin(np.clip(g[2], -1.0, 1.0)))
    mask[np.abs(b) < size] = 0
    return mask
