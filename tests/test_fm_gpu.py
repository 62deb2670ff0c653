"""GPU parity tests of the FM path, through the C ABI, against the oracle.

Tolerances (float path): no multipath filter: max |d| <= 2e-5, rms <= 5e-6 of
full scale over the whole stream (including the start-up transient, because the per-call
schedule is reproduced exactly); identical per-call output lengths; identical stereo
switch-over block.
"""
import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run

pytestmark = pytest.mark.gpu

TOL_MAX, TOL_RMS = 2e-5, 5e-6


def _chunks(n_blocks, per_call):
    o = 0
    while o < n_blocks:
        k = min(per_call, n_blocks - o)
        yield o, k
        o += k


def _run_gpu(dec, iq, blk, blocks_per_call):
    n_blocks = iq.shape[1] // blk
    outs, lens = [], []
    for o, k in _chunks(n_blocks, blocks_per_call):
        a, l = dec.process_blocks(iq[:, o * blk:(o + k) * blk], [blk] * k)
        outs.append(a)
        lens.append(l)
    return np.concatenate(outs, axis=1), np.concatenate(lens)


def test_cfg1_mono_1msps():
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 1.0e6, 2048, 600
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 0, mono=True)[None, :]
    dec = FmDecoder(stereo=False, input_rate=fs, n_channels=1, max_samples_per_call=blk * 256)
    audio, lens = _run_gpu(dec, iq, blk, 256)
    ref_audio, ref_lens, td, st = oracle_fm_run(iq[0], fs, blk, stereo=False, taps=("if",))
    assert list(lens) == list(ref_lens)
    d = audio[0] - ref_audio
    print("cfg1 max", np.abs(d).max(), "rms", np.sqrt(np.mean(d * d)), "n", len(d))
    assert np.abs(d).max() <= TOL_MAX and np.sqrt(np.mean(d * d)) <= TOL_RMS
    s = dec.stats(0)
    assert abs(s.if_rms - st.if_rms) < 1e-5
    assert abs(s.baseband_level - st.baseband_level) < 1e-5
    assert abs(s.if_agc_gain - st.agc_gain) < 1e-4 * st.agc_gain


def test_if_stage_10msps():
    """Stage-level check of the IF resampler (Fs/4 shift + 3 half-bands + 2307-tap LPF + bank)."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 1.0e7, 2048, 200
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 3)[None, :]
    for fs4 in (False, True):
        dec = FmDecoder(stereo=True, input_rate=fs, fs4_shift=fs4, n_channels=1, max_samples_per_call=blk * nblk)
        dec.process_blocks(iq, [blk] * nblk)
        got = dec.tap_if(0)
        _, _, td, _ = oracle_fm_run(iq[0], fs, blk, stereo=True, fs4=fs4, taps=("if",))
        want = np.concatenate(td["if"])
        assert len(got) == len(want)
        d = np.abs(got - want)
        print("IF 10M fs4=%d: n=%d max %.3e (signal rms %.3f)" % (fs4, len(d), d.max(), np.sqrt(np.mean(np.abs(want) ** 2))))
        assert d.max() < 2e-6


@pytest.mark.parametrize("blocks_per_call", [512, 77])
def test_cfg2_stereo_10msps(blocks_per_call):
    from airspy_fmradion_b200 import FmDecoder
    fs, blk = 1.0e7, 2048
    nblk = 3400  # 0.7 s: past the 0.5 s stereo lock
    C = 2
    iq = np.stack([siggen.fm_stereo_iq(fs, blk * nblk, c) for c in range(C)])
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=C, max_samples_per_call=blk * 512)
    audio, lens = _run_gpu(dec, iq, blk, blocks_per_call)
    for c in range(C):
        ref_audio, ref_lens, _, st = oracle_fm_run(iq[c], fs, blk, stereo=True, taps=("if",))
        assert list(lens) == list(ref_lens)
        d = audio[c] - ref_audio
        print("cfg2 ch%d max %.3e rms %.3e n %d" % (c, np.abs(d).max(), np.sqrt(np.mean(d * d)), len(d)))
        assert np.abs(d).max() <= TOL_MAX and np.sqrt(np.mean(d * d)) <= TOL_RMS
        s = dec.stats(c)
        assert s.stereo_detected == st.stereo_detected == 1
        assert s.pll_lock_cnt == st.pll_lock_cnt
        assert abs(s.pilot_level - st.pilot_level) < 1e-5
        # stereo separation is real: L != R in the tail
        tail = audio[c][-2000:]
        assert np.abs(tail[0::2] - tail[1::2]).max() > 0.1


def test_fmfilter_and_pilot_shift_384k():
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 384000.0, 2048, 150
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 1)[None, :]
    for kw in (dict(filter=1), dict(filter=2), dict(pilot_shift=True), dict(deemphasis_us=75.0)):
        dec = FmDecoder(fmfilter=kw.get("filter", 0), stereo=True, deemphasis=kw.get("deemphasis_us", 50.0),
                        pilot_shift=kw.get("pilot_shift", False), input_rate=fs, n_channels=1,
                        max_samples_per_call=blk * nblk)
        audio, lens = dec.process_blocks(iq, [blk] * nblk)
        ref_audio, ref_lens = oracle_fm_run(iq[0], fs, blk, stereo=True, **kw)
        assert list(lens) == list(ref_lens)
        d = audio[0] - ref_audio
        print(kw, "max %.3e rms %.3e" % (np.abs(d).max(), np.sqrt(np.mean(d * d))))
        assert np.abs(d).max() <= TOL_MAX and np.sqrt(np.mean(d * d)) <= TOL_RMS


def test_cfg3_multipath_E200():
    """10 Msps stereo with a static 20 us echo and -E 200 (801-tap NLMS/CMA). Tolerance for the
    adaptive path: max 1e-4, rms 2e-5, coefficient vector relative L2 <= 1e-3."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 1.0e7, 2048, 1500
    echo = (200, 0.3 * np.exp(1j * 0.7))
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 0, echo=echo)[None, :]
    dec = FmDecoder(stereo=True, multipath_stages=200, input_rate=fs, n_channels=1, max_samples_per_call=blk * 500)
    audio, lens = _run_gpu(dec, iq, blk, 500)
    c = None
    from oracle import ref
    if ref.available():
        c = ref.RefChain("fm", fs, stereo=True, mpf_stages=200)
        ref_audio, ref_lens, _ = c.run(iq[0], blk)
        ref_coef = c.mpf_coeffs()
        ref_err = c.stats().mpf_error
    else:
        ref_audio, ref_lens, _, st = oracle_fm_run(iq[0], fs, blk, stereo=True, mpf_stages=200, taps=("if",))
        ref_coef, ref_err = st.mpf_coeffs, st.mpf_error
    assert list(lens) == list(ref_lens)
    d = audio[0] - ref_audio
    coef = dec.get_multipath_coefficients(0)
    rel = np.linalg.norm(coef - ref_coef) / np.linalg.norm(ref_coef)
    print("cfg3 max %.3e rms %.3e coef rel L2 %.3e  err gpu %.4e ref %.4e  |coef-delta| %.3f" % (
        np.abs(d).max(), np.sqrt(np.mean(d * d)), rel, dec.get_multipath_error(0), ref_err,
        np.linalg.norm(ref_coef) - 1))
    assert np.abs(d).max() <= 1e-4 and np.sqrt(np.mean(d * d)) <= 2e-5
    assert rel <= 1e-3
    # the filter really adapted (it is not the identity any more)
    assert np.abs(ref_coef).sum() > 1.05


def test_int16_ingest():
    """IQ delivered as int16 pairs (what FileSource hands over for 16-bit WAV files after
    sf_read_float: value/32768, FileSource.cpp:491-531); conversion fused into the first kernel."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 1.0e7, 2048, 600
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 4)
    q = np.empty((1, len(iq), 2), dtype=np.int16)
    q[0, :, 0] = np.clip(np.round(iq.real * 32768.0), -32768, 32767)
    q[0, :, 1] = np.clip(np.round(iq.imag * 32768.0), -32768, 32767)
    as_float = (q[0, :, 0].astype(np.float32) / np.float32(32768.0)) + 1j * (q[0, :, 1].astype(np.float32) / np.float32(32768.0))
    dec = FmDecoder(stereo=True, input_rate=fs, fs4_shift=True, n_channels=1, max_samples_per_call=blk * 300)
    outs, lens = [], []
    for o in range(0, nblk, 300):
        a, l = dec.process_blocks_i16(q[:, o * blk:(o + 300) * blk], [blk] * 300)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    ref_audio, ref_lens = oracle_fm_run(as_float.astype(np.complex64), fs, blk, stereo=True, fs4=True)
    assert list(lens) == list(ref_lens)
    d = audio[0] - ref_audio
    print("int16 ingest: n=%d max %.3e" % (len(d), np.abs(d).max()))
    assert len(d) > 1000 and np.abs(d).max() <= TOL_MAX
