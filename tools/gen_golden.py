#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz from the COMPILED REFERENCE (oracle/_ref/libfmref.so).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile).
Inputs are not stored: they are regenerated from oracle/siggen.py (deterministic); a CRC of
each input is stored so a generator drift is detected instead of mis-reported as a parity
failure. Outputs are the reference's own audio / taps / statistics.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, siggen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")

CASES = {
    # name: (mode, fs, n_blocks, blk, siggen kwargs, decoder kwargs)
    "fm_mono_1M": ("fm", 1.0e6, 200, 2048, dict(channel=0, mono=True), dict(stereo=False)),
    "fm_stereo_1M": ("fm", 1.0e6, 700, 2048, dict(channel=1), dict(stereo=True)),
    "fm_stereo_10M": ("fm", 1.0e7, 500, 2048, dict(channel=2), dict(stereo=True)),
    "fm_stereo_384k_fs4_narrow": ("fm", 384000.0, 120, 2048, dict(channel=3), dict(stereo=True, fs4=True, filter=2)),
    "fm_stereo_384k_E8": ("fm", 384000.0, 160, 2048, dict(channel=4, echo=(8, 0.3 * np.exp(0.7j))),
                          dict(stereo=True, mpf_stages=8)),
    "fm_stereo_384k_blk777": ("fm", 384000.0, 300, 777, dict(channel=5), dict(stereo=True)),
    "am_384k": ("am", 384000.0, 150, 2048, dict(channel=0), dict()),
}


def make_input(mode, fs, n, kw):
    if mode == "fm":
        return siggen.fm_stereo_iq(fs, n, **kw)
    return siggen.am_iq(fs, n, **kw)


def window(a, k=1500):
    """first k, last k samples (the whole array when it is short)."""
    if len(a) <= 2 * k:
        return a.copy(), np.array([0, len(a)])
    return np.concatenate([a[:k], a[-k:]]), np.array([k, len(a)])


def main():
    assert ref.available(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    out = {}
    for name, (mode, fs, nblk, blk, skw, dkw) in CASES.items():
        iq = make_input(mode, fs, nblk * blk, skw)
        out[name + "/crc"] = np.array([zlib.crc32(iq.tobytes())], dtype=np.uint32)
        c = ref.RefChain(mode, fs, **dkw)
        audio, lens, td = c.run(iq, blk, taps=("if",) if fs > 384000 else ())
        st = c.stats()
        out[name + "/lens"] = lens.astype(np.int32)
        a, w = window(audio)
        out[name + "/audio"] = a
        out[name + "/audio_window"] = w
        out[name + "/audio_sum"] = np.array([audio.sum(), np.abs(audio).sum()])
        if td.get("if"):
            ifs = np.concatenate(td["if"])
            a, w = window(ifs, 1000)
            out[name + "/if"] = a
            out[name + "/if_window"] = w
        if mode == "fm":
            out[name + "/stats"] = np.array([st.stereo_detected, st.tuning_offset, st.baseband_level, st.pilot_level,
                                             st.if_rms, st.mpf_error, st.agc_gain, st.pll_freq, st.pll_phase,
                                             st.pll_lock_cnt, st.decoder_calls], dtype=np.float64)
            if dkw.get("mpf_stages"):
                out[name + "/mpf_coeffs"] = c.mpf_coeffs()
        else:
            out[name + "/stats"] = np.array([st.baseband_level, st.af_agc_gain, st.if_agc_gain, st.if_rms,
                                             st.decoder_calls], dtype=np.float64)
        c.close()
        print(name, "audio", len(audio), "calls", len(lens))
    # resampler stage vectors: seeded noise through each shipped chain (one real lane)
    for (src, dst, kind, n) in [(1e7, 384000.0, 0, 150000), (6e6, 384000.0, 0, 120000), (2.5e6, 384000.0, 0, 60000),
                                (1e6, 384000.0, 0, 40000), (384000.0, 48000.0, 0, 50000), (384000.0, 48000.0, 1, 50000)]:
        x = np.random.Generator(np.random.PCG64(99)).standard_normal(n)
        r = ref.R8b(src, dst, kind)
        ys, ls = [], []
        for o in range(0, n, 3000):
            y = r.process(x[o:o + 3000])
            ys.append(y)
            ls.append(len(y))
        key = "r8b_%d_%d_%d" % (src, dst, kind)
        out[key + "/crc"] = np.array([zlib.crc32(x.tobytes())], dtype=np.uint32)
        out[key + "/out"] = np.concatenate(ys)
        out[key + "/lens"] = np.array(ls, dtype=np.int32)
    # fast_atan2f known answers on a grid
    g = np.linspace(-1.5, 1.5, 41, dtype=np.float32)
    out["fast_atan2f/grid"] = g
    out["fast_atan2f/val"] = np.array([[ref.fast_atan2f(y, x) for x in g] for y in g], dtype=np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
