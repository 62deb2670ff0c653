"""CPU tests of the oracle for the DSB / USB / LSB / CW / WSPR branches of AmDecoder::process: the
plain-C restatement against golden vectors made from the compiled reference
(tools/gen_golden_ammodes.py) and, when oracle/_ref was built here, against it live."""
import os
import zlib

import numpy as np
import pytest

from oracle import ref, restate, siggen

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden_ammodes_v1.npz")

# must mirror tools/gen_golden_ammodes.py CASES: (modtype, fs, n_blocks, blk, channel)
CASES = {
    "dsb_384k": (3, 384000.0, 100, 2048, 0),
    "usb_384k": (4, 384000.0, 100, 2048, 1),
    "lsb_48k": (5, 48000.0, 40, 1000, 2),
    "cw_384k": (6, 384000.0, 100, 2048, 3),
    "wspr_48k_blk777": (7, 48000.0, 60, 777, 4),
}


def case_input(name, g):
    mt, fs, nblk, blk, ch = CASES[name]
    iq = siggen.ssb_iq(fs, nblk * blk, ch)
    if zlib.crc32(iq.tobytes()) != int(g[name + "/crc"][0]):
        pytest.skip("synthetic generator output differs from the one the golden vectors were made with")
    return iq


def window_err(g, name, full):
    w, want = g[name + "/audio_window"], g[name + "/audio"]
    k, n = int(w[0]), int(w[1])
    assert len(full) == n
    got = full if k == 0 else np.concatenate([full[:k], full[-k:]])
    return float(np.abs(got - want).max())


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_ammodes_vs_golden(name):
    g = np.load(GOLDEN)
    mt, fs, nblk, blk, ch = CASES[name]
    iq = case_input(name, g)
    audio, lens, _, st = restate.am_run(iq, fs, blk, modtype=mt)
    assert list(lens) == list(g[name + "/lens"])
    # 2049-tap float FIR: the reference build vectorises its summation, 1e-6 covers the order
    assert window_err(g, name, audio) <= 1e-6
    s = g[name + "/stats"]
    assert abs(st.baseband_level - s[0]) < 1e-5 and abs(st.af_agc_gain - s[1]) < 1e-5
    assert abs(st.if_agc_gain - s[2]) < 1e-4 * s[2] and abs(st.if_rms - s[3]) < 1e-6
    assert st.decoder_calls == int(s[4])


def test_usb_selects_the_upper_sideband():
    """Sanity of the signal path itself (not only parity): USB keeps the +700/+1900 Hz tones and
    rejects the -1100 Hz one; LSB does the opposite."""
    fs, blk, nblk = 48000.0, 1000, 60
    iq = siggen.ssb_iq(fs, blk * nblk, 0)
    spec = {}
    for mt in (4, 5):
        a, _, _, _ = restate.am_run(iq, fs, blk, modtype=mt)
        x = a[20000:20000 + 32768]
        spec[mt] = np.abs(np.fft.rfft(x * np.hanning(len(x))))
    f = np.fft.rfftfreq(32768, 1 / fs)

    def peak(sp, hz):
        return sp[np.abs(f - hz) < 15].max()
    assert peak(spec[4], 700) > 30 * peak(spec[4], 1100)
    assert peak(spec[5], 1100) > 30 * peak(spec[5], 700)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libfmref.so not built in this environment")
@pytest.mark.parametrize("mt", [3, 4, 5, 6, 7])
def test_restatement_ammodes_vs_compiled_reference_live(mt):
    fs, blk = 384000.0, 3000
    iq = siggen.ssb_iq(fs, blk * 60, 6)
    a, la, _, sa = restate.am_run(iq, fs, blk, modtype=mt)
    c = ref.RefChain("am", fs, modtype=mt)
    b, lb, _ = c.run(iq, blk)
    sb = c.stats()
    assert list(la) == list(lb) and len(a) > 15000
    assert np.abs(a - b).max() < 1e-6
    assert abs(sa.if_rms - sb.if_rms) < 1e-6 and abs(sa.af_agc_gain - sb.af_agc_gain) < 1e-6
