#!/usr/bin/env python
"""Golden vectors for the NBFM path (NbfmDecoder::process behind FourthConverterIQ + IfResampler),
made from the compiled reference (oracle/_ref/libfmref.so) like tools/gen_golden.py does for FM/AM;
kept in their own file so that the older fixtures stay byte-identical.

    python tools/gen_golden_nbfm.py        # build container only (needs /root/reference)
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, siggen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_nbfm_v1.npz")

CASES = {
    # name: (fs, n_blocks, blk, siggen kwargs, decoder kwargs)
    "nbfm_384k": (384000.0, 150, 2048, dict(channel=0), dict(filter=0)),
    "nbfm_384k_narrow_fs4": (384000.0, 120, 2048, dict(channel=1, dev=2000.0), dict(filter=2, fs4=True)),
    "nbfm_48k_wide17k": (48000.0, 60, 1024, dict(channel=2, dev=9000.0), dict(filter=3, freq_dev=17000.0)),
    "nbfm_48k_blk333": (48000.0, 150, 333, dict(channel=3), dict(filter=1)),
}


def window(a, k=1500):
    if len(a) <= 2 * k:
        return a.copy(), np.array([0, len(a)])
    return np.concatenate([a[:k], a[-k:]]), np.array([k, len(a)])


def main():
    assert ref.available(), "build oracle/_ref first"
    out = {}
    for name, (fs, nblk, blk, skw, dkw) in CASES.items():
        iq = siggen.nbfm_iq(fs, nblk * blk, **skw)
        out[name + "/crc"] = np.array([zlib.crc32(iq.tobytes())], dtype=np.uint32)
        c = ref.RefChain("nbfm", fs, **dkw)
        audio, lens, _ = c.run(iq, blk)
        st = c.stats()
        out[name + "/lens"] = lens.astype(np.int32)
        a, w = window(audio)
        out[name + "/audio"] = a
        out[name + "/audio_window"] = w
        out[name + "/audio_sum"] = np.array([audio.sum(), np.abs(audio).sum()])
        out[name + "/stats"] = np.array([st.tuning_offset, st.baseband_level, st.if_rms, st.if_agc_gain,
                                         st.decoder_calls], dtype=np.float64)
        c.close()
        print(name, "audio", len(audio), "calls", len(lens), "tuning offset %.2f Hz" % st.tuning_offset)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
