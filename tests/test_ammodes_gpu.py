"""GPU parity of the DSB / USB / LSB / CW / WSPR branches of AmDecoder::process through the C ABI:
against the oracle on seeded inputs and against the committed golden vectors."""
import numpy as np
import pytest

from oracle import ref, restate, siggen
from tests.test_ammodes_cpu import CASES, GOLDEN, case_input, window_err

pytestmark = pytest.mark.gpu

TOL_MAX, TOL_RMS = 2e-5, 5e-6


def _oracle(iq, fs, blk, mt):
    if ref.available():
        c = ref.RefChain("am", fs, modtype=mt)
        a, l, _ = c.run(iq, blk)
        st = c.stats()
        c.close()
        return a, l, st
    a, l, _, st = restate.am_run(iq, fs, blk, modtype=mt)
    return a, l, st


@pytest.mark.parametrize("mt,fs", [(3, 384000.0), (4, 384000.0), (5, 48000.0), (6, 384000.0), (7, 48000.0)])
def test_ammodes_multichannel(mt, fs):
    from airspy_fmradion_b200 import AmDecoder
    blk, nblk, C = 2048, 100 if fs > 48000 else 30, 5
    iq = np.stack([siggen.ssb_iq(fs, blk * nblk, c) for c in range(C)])
    dec = AmDecoder(mode=mt, input_rate=fs, n_channels=C, max_samples_per_call=blk * 40)
    outs, lens = [], []
    for o in range(0, nblk, 40):   # several calls: the FineTuner indices carry over
        k = min(40, nblk - o)
        a, l = dec.process_blocks(iq[:, o * blk:(o + k) * blk], [blk] * k)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    for c in (0, C - 1):
        want, want_lens, st = _oracle(iq[c], fs, blk, mt)
        assert list(lens) == list(want_lens)
        d = audio[c] - want
        print("mode %d fs=%g ch%d: n=%d max %.3e rms %.3e (audio rms %.3f)" % (
            mt, fs, c, len(d), np.abs(d).max(), np.sqrt(np.mean(d * d)), np.sqrt(np.mean(want * want))))
        assert len(d) > 10000 and np.abs(d).max() <= TOL_MAX and np.sqrt(np.mean(d * d)) <= TOL_RMS
        s = dec.stats(c)
        assert abs(s.if_rms - st.if_rms) < 1e-5 and abs(s.baseband_level - st.baseband_level) < 1e-5
        assert abs(s.af_agc_gain - st.af_agc_gain) < 1e-4 and abs(s.if_agc_gain - st.if_agc_gain) < 1e-3 * st.if_agc_gain


@pytest.mark.parametrize("name", sorted(CASES))
def test_ammodes_gpu_vs_golden(name):
    from airspy_fmradion_b200 import AmDecoder
    g = np.load(GOLDEN)
    mt, fs, nblk, blk, ch = CASES[name]
    iq = case_input(name, g)[None, :]
    dec = AmDecoder(mode=mt, input_rate=fs, n_channels=1, max_samples_per_call=nblk * blk, max_blocks_per_call=nblk)
    audio, lens = dec.process_blocks(iq, [blk] * nblk)
    assert list(lens) == list(g[name + "/lens"])
    err = window_err(g, name, audio[0])
    print(name, "max |gpu - golden| =", err)
    assert err <= TOL_MAX
    np.testing.assert_allclose(np.abs(audio[0]).sum(), g[name + "/audio_sum"][1], rtol=1e-5)
    s, st = g[name + "/stats"], dec.stats(0)
    assert abs(st.baseband_level - s[0]) < 1e-5 and abs(st.if_rms - s[3]) < 1e-5 and st.decoder_calls == int(s[4])
