"""Stand-in module named ``pys2let`` (see shims/pyssht.py).  TEST INFRASTRUCTURE ONLY."""
from oracle import s2let_ref as _w


def mw_size(L):
    return _w.mw_size(L)


def pys2let_j_max(B, L, J_min):
    return _w.j_max(L, B)


def wavelet_tiling(B, L, N, J_min, spin):
    return _w.wavelet_tiling(B, L, N, J_min, spin)


def analysis_px2wav(f, B, L, J_min, N, spin, upsample=0):
    return _w.analysis_px2wav(f, B, L, J_min, N, spin, upsample)


def synthesis_wav2px(f_wav, f_scal, B, L, J_min, N, spin, upsample=0):
    return _w.synthesis_wav2px(f_wav, f_scal, B, L, J_min, N, spin, upsample)


def analysis_adjoint_wav2px(f_wav, f_scal, B, L, J_min, N, spin, upsample=0):
    return _w.analysis_adjoint_wav2px(f_wav, f_scal, B, L, J_min, N, spin, upsample)


def synthesis_adjoint_px2wav(f, B, L, J_min, N, spin, upsample=0):
    return _w.synthesis_adjoint_px2wav(f, B, L, J_min, N, spin, upsample)


def alm2map_mw(flm, L, spin):
    return _w.alm2map_mw(flm, L, spin)


def lm_hp2lm(alm, L):
    return _w.lm_hp2lm(alm, L)
