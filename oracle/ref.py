"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libfmref.so.

libfmref.so is the reference's own DSP classes (FourthConverterIQ, IfResampler, FmDecoder,
AmDecoder, r8brain) compiled unmodified from /root/reference by oracle/Makefile. The
built library travels to the GPU box with the snapshot; /root/reference itself is never
read at run time.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfmref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_fm_create.restype = C.c_void_p
        L.ref_fm_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint]
        L.ref_am_create.restype = C.c_void_p
        L.ref_am_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        L.ref_nbfm_create.restype = C.c_void_p
        L.ref_nbfm_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double]
        L.ref_nbfm_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_process_block.restype = C.c_int
        L.ref_process_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for name in ("ref_tap_if", "ref_tap_fm_preDisc", "ref_tap_fm_mpx", "ref_tap_fm_mono384",
                     "ref_tap_fm_stereo384", "ref_tap_fm_mono48_first", "ref_tap_fm_stereo48_first",
                     "ref_fm_mpf_coeffs", "ref_fm_pps"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_fm_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_am_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_r8b_create.restype = C.c_void_p
        L.ref_r8b_create.argtypes = [C.c_double, C.c_double, C.c_int]
        L.ref_r8b_destroy.argtypes = [C.c_void_p]
        L.ref_r8b_process.restype = C.c_int
        L.ref_r8b_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_r8b_dump.restype = C.c_int
        L.ref_r8b_dump.argtypes = [C.c_double, C.c_double, C.c_int, C.c_char_p]
        L.ref_filter_table.restype = C.c_int
        L.ref_filter_table.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.ref_fast_atan2f.restype = C.c_float
        L.ref_fast_atan2f.argtypes = [C.c_float, C.c_float]
        L.ref_bench.restype = C.c_double
        L.ref_bench.argtypes = [C.c_int, C.c_double, C.c_int, C.c_uint, C.c_int, C.c_long, C.c_int,
                                C.c_void_p, C.c_long]
        _lib = L
    return _lib


class FmStats(C.Structure):
    _fields_ = [("stereo_detected", C.c_int), ("tuning_offset", C.c_float),
                ("baseband_level", C.c_float), ("pilot_level", C.c_double),
                ("if_rms", C.c_float), ("mpf_error", C.c_double), ("agc_gain", C.c_float),
                ("pll_freq", C.c_double), ("pll_phase", C.c_double), ("pll_lock_cnt", C.c_int),
                ("decoder_calls", C.c_uint64), ("n_pps", C.c_int)]


class AmStats(C.Structure):
    _fields_ = [("baseband_level", C.c_double), ("af_agc_gain", C.c_float),
                ("if_agc_gain", C.c_float), ("if_rms", C.c_float), ("decoder_calls", C.c_uint64)]


class NbfmStats(C.Structure):
    _fields_ = [("tuning_offset", C.c_float), ("baseband_level", C.c_float), ("if_rms", C.c_float),
                ("if_agc_gain", C.c_float), ("decoder_calls", C.c_uint64)]


MODTYPE_AM = 2  # include/SoftFM.h:49 enum class ModType { FM, NBFM, AM, ... }
MODTYPE_NBFM = 1


class RefChain:
    """One reference receive chain (front end + decoder) fed block by block."""

    def __init__(self, mode="fm", ifrate=384000.0, fs4=False, filter=0, stereo=True,
                 deemphasis_us=50.0, pilot_shift=False, mpf_stages=0, freq_dev=8000.0, modtype=MODTYPE_AM):
        L = lib()
        self.mode = mode
        if mode == "fm":
            self.h = L.ref_fm_create(ifrate, int(fs4), filter, int(stereo), deemphasis_us,
                                     int(pilot_shift), mpf_stages)
        elif mode == "nbfm":
            self.h = L.ref_nbfm_create(ifrate, int(fs4), filter, float(freq_dev))
        else:
            self.h = L.ref_am_create(ifrate, int(fs4), filter, int(modtype))
        self._audio = np.empty(1 << 18, dtype=np.float64)

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process_block(self, iq):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        m = lib().ref_process_block(self.h, iq.ctypes.data, len(iq), self._audio.ctypes.data,
                                    len(self._audio))
        assert m >= 0
        return self._audio[:m].copy()

    def _tap(self, name, dtype, width=1, cap=1 << 17):
        buf = np.empty(cap * width, dtype=dtype)
        n = getattr(lib(), name)(self.h, buf.ctypes.data, cap)
        assert n <= cap
        return buf[:n * width].copy()

    def tap_if(self):
        return self._tap("ref_tap_if", np.float32, 2).view(np.complex64)

    def tap_predisc(self):
        return self._tap("ref_tap_fm_preDisc", np.float32, 2).view(np.complex64)

    def tap_mpx(self):
        return self._tap("ref_tap_fm_mpx", np.float32)

    def tap_mono384(self):
        return self._tap("ref_tap_fm_mono384", np.float64)

    def tap_stereo384(self):
        return self._tap("ref_tap_fm_stereo384", np.float64)

    def tap_mono48_first(self):
        return self._tap("ref_tap_fm_mono48_first", np.float64)

    def tap_stereo48_first(self):
        return self._tap("ref_tap_fm_stereo48_first", np.float64)

    def mpf_coeffs(self):
        return self._tap("ref_fm_mpf_coeffs", np.float32, 2, cap=4096).view(np.complex64)

    def pps(self):
        return self._tap("ref_fm_pps", np.float64, 3, cap=64).reshape(-1, 3)

    def stats(self):
        if self.mode == "fm":
            s = FmStats()
            lib().ref_fm_stats(self.h, C.byref(s))
        elif self.mode == "nbfm":
            s = NbfmStats()
            lib().ref_nbfm_stats(self.h, C.byref(s))
        else:
            s = AmStats()
            lib().ref_am_stats(self.h, C.byref(s))
        return s

    def run(self, iq, blklen=2048, taps=()):
        """Feed `iq` in blocks of blklen; return (audio_concat, per_call_len, tapdict)."""
        out, lens = [], []
        td = {k: [] for k in taps}
        for o in range(0, len(iq), blklen):
            a = self.process_block(iq[o:o + blklen])
            out.append(a)
            lens.append(len(a))
            for k in taps:
                td[k].append(getattr(self, "tap_" + k)())
        audio = np.concatenate(out) if out else np.empty(0)
        return audio, np.array(lens, dtype=np.int64), td


def filter_table(name):
    buf = np.empty(4096, dtype=np.float64)
    n = lib().ref_filter_table(name.encode(), buf.ctypes.data, len(buf))
    assert n > 0, name
    return buf[:n].copy()


def fast_atan_table():
    buf = np.empty(512, dtype=np.float32)
    L = lib()
    L.ref_fast_atan_table.restype = C.c_int
    L.ref_fast_atan_table.argtypes = [C.c_void_p, C.c_int]
    n = L.ref_fast_atan_table(buf.ctypes.data, len(buf))
    return buf[:n].copy()


def fast_atan2f(y, x):
    return float(lib().ref_fast_atan2f(float(y), float(x)))


def r8b_dump(src, dst, kind):
    """Return the stage list of the r8brain chain for (src, dst); kind 0 = IF (24-bit spec),
    1 = audio (default spec)."""
    import tempfile
    with tempfile.NamedTemporaryFile("r", suffix=".txt") as tf:
        rc = lib().ref_r8b_dump(src, dst, kind, tf.name.encode())
        assert rc == 0, rc
        lines = open(tf.name).read().split("\n")
    it = iter(lines)
    hdr = next(it).split()
    assert hdr[0] == "chain"
    stages = []
    for _ in range(int(hdr[1])):
        h = next(it).split()
        if h[0] == "hb":
            n = int(h[1])
            stages.append(dict(type="hb", taps=[float(next(it)) for _ in range(n)], latency=int(h[2])))
        elif h[0] == "bc":
            klen = int(h[1])
            stages.append(dict(type="bc", klen=klen, inputlen=int(h[2]), latency=int(h[3]), up=int(h[4]),
                               down=int(h[5]), outoffset=int(h[6]), inputdelay=int(h[7]),
                               downskipinit=int(h[8]), taps=[float(next(it)) for _ in range(klen)]))
        elif h[0] == "fi":
            ins, outs, fl = int(h[1]), int(h[2]), int(h[3])
            stages.append(dict(type="fi", instep=ins, outstep=outs, flen=fl, initfracposw=int(h[4]),
                               latency=int(h[5]), taps=[float(next(it)) for _ in range(outs * fl)]))
    return stages


class R8b:
    def __init__(self, src, dst, kind):
        self.h = lib().ref_r8b_create(src, dst, kind)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x) + 64, dtype=np.float64)
        m = lib().ref_r8b_process(self.h, x.ctypes.data, len(x), out.ctypes.data, len(out))
        assert m >= 0
        return out[:m].copy()

    def __del__(self):
        try:
            lib().ref_r8b_destroy(self.h)
        except Exception:
            pass


def bench(mode, ifrate, stereo, mpf_stages, nthreads, blocks, blklen, iq):
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    return lib().ref_bench(mode, ifrate, int(stereo), mpf_stages, nthreads, blocks, blklen,
                           iq.ctypes.data, len(iq))
