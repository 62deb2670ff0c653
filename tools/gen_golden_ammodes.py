#!/usr/bin/env python
"""Golden vectors for the DSB / USB / LSB / CW / WSPR branches of AmDecoder::process
(AmDecode.cpp:96-218, FineTuner.cpp:55-70), made from the compiled reference; own file so that the
older fixtures stay byte-identical.

    python tools/gen_golden_ammodes.py        # build container only (needs /root/reference)
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref, siggen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_ammodes_v1.npz")

MODES = {"dsb": 3, "usb": 4, "lsb": 5, "cw": 6, "wspr": 7}
CASES = {
    # name: (modtype, fs, n_blocks, blk, channel)
    "dsb_384k": (3, 384000.0, 100, 2048, 0),
    "usb_384k": (4, 384000.0, 100, 2048, 1),
    "lsb_48k": (5, 48000.0, 40, 1000, 2),
    "cw_384k": (6, 384000.0, 100, 2048, 3),
    "wspr_48k_blk777": (7, 48000.0, 60, 777, 4),
}


def window(a, k=1500):
    if len(a) <= 2 * k:
        return a.copy(), np.array([0, len(a)])
    return np.concatenate([a[:k], a[-k:]]), np.array([k, len(a)])


def main():
    assert ref.available(), "build oracle/_ref first"
    out = {}
    for name, (mt, fs, nblk, blk, ch) in CASES.items():
        iq = siggen.ssb_iq(fs, nblk * blk, ch)
        out[name + "/crc"] = np.array([zlib.crc32(iq.tobytes())], dtype=np.uint32)
        c = ref.RefChain("am", fs, modtype=mt)
        audio, lens, _ = c.run(iq, blk)
        st = c.stats()
        out[name + "/lens"] = lens.astype(np.int32)
        a, w = window(audio)
        out[name + "/audio"] = a
        out[name + "/audio_window"] = w
        out[name + "/audio_sum"] = np.array([audio.sum(), np.abs(audio).sum()])
        out[name + "/stats"] = np.array([st.baseband_level, st.af_agc_gain, st.if_agc_gain, st.if_rms,
                                         st.decoder_calls], dtype=np.float64)
        c.close()
        print(name, "audio", len(audio), "rms %.3f" % np.sqrt(np.mean(audio ** 2)))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
