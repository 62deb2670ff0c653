"""CPU tests of the NBFM oracle: the plain-C restatement (oracle/restate, NbfmDecoder::process)
against the golden vectors made from the compiled reference (tools/gen_golden_nbfm.py) and, when
oracle/_ref was built here, against the compiled reference live."""
import os
import zlib

import numpy as np
import pytest

from oracle import ref, restate, siggen

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden_nbfm_v1.npz")

# must mirror tools/gen_golden_nbfm.py CASES
CASES = {
    "nbfm_384k": (384000.0, 150, 2048, dict(channel=0), dict(filter=0)),
    "nbfm_384k_narrow_fs4": (384000.0, 120, 2048, dict(channel=1, dev=2000.0), dict(filter=2, fs4=True)),
    "nbfm_48k_wide17k": (48000.0, 60, 1024, dict(channel=2, dev=9000.0), dict(filter=3, freq_dev=17000.0)),
    "nbfm_48k_blk333": (48000.0, 150, 333, dict(channel=3), dict(filter=1)),
}


def case_input(name, g):
    fs, nblk, blk, skw, dkw = CASES[name]
    iq = siggen.nbfm_iq(fs, nblk * blk, **skw)
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
def test_restatement_nbfm_vs_golden(name):
    g = np.load(GOLDEN)
    fs, nblk, blk, skw, dkw = CASES[name]
    iq = case_input(name, g)
    audio, lens, _, st = restate.nbfm_run(iq, fs, blk, **dkw)
    assert list(lens) == list(g[name + "/lens"])
    assert window_err(g, name, audio) <= 1e-7
    np.testing.assert_allclose([audio.sum(), np.abs(audio).sum()], g[name + "/audio_sum"], rtol=1e-6, atol=1e-6)
    s = g[name + "/stats"]
    assert abs(st.tuning_offset - s[0]) < 1e-2 and abs(st.baseband_level - s[1]) < 1e-6
    assert abs(st.if_rms - s[2]) < 1e-6 and abs(st.if_agc_gain - s[3]) < 1e-4 * s[3]
    assert st.decoder_calls == int(s[4])
    # the signal really is demodulated: the two message tones dominate the audio spectrum
    if name == "nbfm_384k":
        a = audio[4800:4800 + 16384]
        spec = np.abs(np.fft.rfft(a * np.hanning(len(a))))
        f = np.fft.rfftfreq(len(a), 1 / 48000.0)
        assert abs(f[np.argmax(spec)] - 600.0) < 10.0


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libfmref.so not built in this environment")
def test_restatement_nbfm_vs_compiled_reference_live():
    fs, blk = 384000.0, 3000
    iq = siggen.nbfm_iq(fs, blk * 90, 5)
    a, la, _, sa = restate.nbfm_run(iq, fs, blk, filter=2, freq_dev=5000.0)
    c = ref.RefChain("nbfm", fs, filter=2, freq_dev=5000.0)
    b, lb, _ = c.run(iq, blk)
    sb = c.stats()
    assert list(la) == list(lb) and len(a) > 20000
    # the float IF filter's summation order is the compiler's (vectorised in the reference build): 1e-7
    assert np.abs(a - b).max() < 1e-7
    assert abs(sa.if_rms - sb.if_rms) < 1e-6 and abs(sa.tuning_offset - sb.tuning_offset) < 1e-3
