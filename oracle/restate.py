"""TEST INFRASTRUCTURE ONLY: ctypes binding of the plain-C restatement (oracle/restate)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "restate")
LIB_PATH = os.path.join(_DIR, "liboracle.so")
_lib = None


class FmStats(C.Structure):
    _fields_ = [("stereo_detected", C.c_int), ("tuning_offset", C.c_float), ("baseband_level", C.c_float),
                ("pilot_level", C.c_double), ("if_rms", C.c_float), ("mpf_error", C.c_double),
                ("agc_gain", C.c_float), ("pll_freq", C.c_double), ("pll_phase", C.c_double),
                ("pll_lock_cnt", C.c_int), ("decoder_calls", C.c_uint64), ("n_pps", C.c_int)]


class AmStats(C.Structure):
    _fields_ = [("baseband_level", C.c_double), ("af_agc_gain", C.c_float), ("if_agc_gain", C.c_float),
                ("if_rms", C.c_float), ("decoder_calls", C.c_uint64)]


class NbfmStats(C.Structure):
    _fields_ = [("tuning_offset", C.c_float), ("baseband_level", C.c_float), ("if_rms", C.c_float),
                ("if_agc_gain", C.c_float), ("decoder_calls", C.c_uint64)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-s", "-C", _DIR])
        L = C.CDLL(LIB_PATH)
        L.orc_fm_create.restype = C.c_void_p
        L.orc_fm_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint]
        L.orc_am_create.restype = C.c_void_p
        L.orc_am_create.argtypes = [C.c_double, C.c_int, C.c_int]
        L.orc_am_create_mode.restype = C.c_void_p
        L.orc_am_create_mode.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_nbfm_create.restype = C.c_void_p
        L.orc_nbfm_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double]
        L.orc_nbfm_stats.argtypes = [C.c_void_p, C.c_void_p]
        for n in ("orc_fm_destroy", "orc_am_destroy", "orc_r8_destroy", "orc_nbfm_destroy"):
            getattr(L, n).argtypes = [C.c_void_p]
            getattr(L, n).restype = None
        for n in ("orc_fm_process_block", "orc_am_process_block", "orc_nbfm_process_block"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for n in ("orc_fm_tap_if", "orc_fm_mpf_coeffs", "orc_fm_pps"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_fm_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_am_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_chain_out.restype = C.c_int64
        L.orc_chain_out.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int64]
        L.orc_r8_create.restype = C.c_void_p
        L.orc_r8_create.argtypes = [C.c_double, C.c_double, C.c_int]
        L.orc_r8_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib = L
    return _lib


class _Stats:
    pass


def _copy_stats(s):
    o = _Stats()
    for f, _ in s._fields_:
        setattr(o, f, getattr(s, f))
    return o


def fm_run(iq, fs, blk, stereo=True, fs4=False, filter=0, deemphasis_us=50.0, pilot_shift=False, mpf_stages=0,
           taps=()):
    L = lib()
    h = L.orc_fm_create(fs, int(fs4), filter, int(stereo), deemphasis_us, int(pilot_shift), mpf_stages)
    assert h, "restatement has no tables for this rate"
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    audio = np.empty(1 << 17, dtype=np.float64)
    out, lens = [], []
    td = {k: [] for k in taps}
    tapbuf = np.empty(2 * (1 << 17), dtype=np.float32)
    for o in range(0, len(iq), blk):
        b = iq[o:o + blk]
        m = L.orc_fm_process_block(h, b.ctypes.data, len(b), audio.ctypes.data, len(audio))
        assert m >= 0
        out.append(audio[:m].copy())
        lens.append(m)
        if "if" in td:
            n = L.orc_fm_tap_if(h, tapbuf.ctypes.data, 1 << 17)
            td["if"].append(tapbuf[:2 * n].copy().view(np.complex64))
    s = FmStats()
    L.orc_fm_stats(h, C.byref(s))
    st = _copy_stats(s)
    cb = np.zeros(2 * 1100, dtype=np.float32)
    n = L.orc_fm_mpf_coeffs(h, cb.ctypes.data, 1100)
    st.mpf_coeffs = cb[:2 * n].copy().view(np.complex64)
    pb = np.zeros(48, dtype=np.float64)
    n = L.orc_fm_pps(h, pb.ctypes.data, 16)
    st.pps = pb[:3 * min(n, 16)].reshape(-1, 3).copy()
    L.orc_fm_destroy(h)
    return (np.concatenate(out) if out else np.empty(0)), np.array(lens, dtype=np.int64), td, st


def am_run(iq, fs, blk, filter=0, fs4=False, modtype=2):
    L = lib()
    h = L.orc_am_create_mode(fs, int(fs4), filter, int(modtype))
    assert h
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    audio = np.empty(1 << 17, dtype=np.float64)
    out, lens = [], []
    for o in range(0, len(iq), blk):
        b = iq[o:o + blk]
        m = L.orc_am_process_block(h, b.ctypes.data, len(b), audio.ctypes.data, len(audio))
        assert m >= 0
        out.append(audio[:m].copy())
        lens.append(m)
    s = AmStats()
    L.orc_am_stats(h, C.byref(s))
    st = _copy_stats(s)
    L.orc_am_destroy(h)
    return (np.concatenate(out) if out else np.empty(0)), np.array(lens, dtype=np.int64), {}, st


def nbfm_run(iq, fs, blk, filter=0, fs4=False, freq_dev=8000.0):
    L = lib()
    h = L.orc_nbfm_create(fs, int(fs4), filter, float(freq_dev))
    assert h
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    audio = np.empty(1 << 17, dtype=np.float64)
    out, lens = [], []
    for o in range(0, len(iq), blk):
        b = iq[o:o + blk]
        m = L.orc_nbfm_process_block(h, b.ctypes.data, len(b), audio.ctypes.data, len(audio))
        assert m >= 0
        out.append(audio[:m].copy())
        lens.append(m)
    s = NbfmStats()
    L.orc_nbfm_stats(h, C.byref(s))
    st = _copy_stats(s)
    L.orc_nbfm_destroy(h)
    return (np.concatenate(out) if out else np.empty(0)), np.array(lens, dtype=np.int64), {}, st


def chain_out(src, dst, kind, n):
    return int(lib().orc_chain_out(src, dst, kind, n))


class R8:
    def __init__(self, src, dst, kind):
        self.h = lib().orc_r8_create(src, dst, kind)
        assert self.h

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x) + 64, dtype=np.float64)
        m = lib().orc_r8_process(self.h, x.ctypes.data, len(x), out.ctypes.data, len(out))
        assert m >= 0
        return out[:m].copy()

    def __del__(self):
        try:
            lib().orc_r8_destroy(self.h)
        except Exception:
            pass
