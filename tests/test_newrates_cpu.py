"""Rate pairs whose tables were added at the end of round 1 (3 M, 2.4 M, 1.44 M, 1.2 M, 1.152 M, 960 k, 912 k, 768 k -> 384 k):
pinned on the CPU against the COMPILED REFERENCE — the plain-C restatement (which reads the same generated tables) must
reproduce the reference's audio and per-call sizes, and the library's host-side schedule must reproduce the per-call
sizes. GPU parity: tests/test_newrates_gpu.py (green on a B200 since round 2, profiles/pytest_newrates_r02.log)."""
import numpy as np
import pytest

from oracle import ref, restate, siggen

RATES = [3.0e6, 2.4e6, 2.048e6, 1.44e6, 1.2e6, 960000.0, 912000.0, 768000.0, 1152000.0]
AM_RATES = [192000.0, 256000.0, 768000.0, 1.0e6]  # -> 48 kHz (AmDecoder / NbfmDecoder handle)
needs_ref = pytest.mark.skipif(not ref.available(), reason="compiled reference (oracle/_ref) not built")


def _blocks(fs):
    # long enough to get past both resamplers' start-up latency (first audio after 29 591 IF samples)
    return int(np.ceil(0.11 * fs / 2048)) + 8


@needs_ref
@pytest.mark.parametrize("fs", RATES)
def test_restatement_and_schedule_match_reference(fs):
    blk, nblk = 2048, _blocks(fs)
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 1)
    c = ref.RefChain("fm", fs, stereo=True)
    ref_audio, ref_lens, _ = c.run(iq, blk)
    c.close()
    audio, lens, _, _ = restate.fm_run(iq, fs, blk, stereo=True)
    assert list(lens) == list(ref_lens)
    assert len(ref_audio) > 500
    assert np.abs(audio - ref_audio).max() <= 1e-9
    from airspy_fmradion_b200 import _capi
    L = _capi.lib()
    bl = np.full(nblk, blk, dtype=np.uint32)
    out = np.zeros(nblk, dtype=np.uint32)
    _capi.check(L.fmr_fm_schedule(fs, 1, 0, bl.ctypes.data, nblk, None, out.ctypes.data))
    assert list(out) == list(ref_lens)


@needs_ref
@pytest.mark.parametrize("fs", AM_RATES)
def test_am_restatement_and_schedule_match_reference(fs):
    blk = 2048
    nblk = int(np.ceil(0.4 * fs / blk)) + 4
    iq = siggen.am_iq(fs, blk * nblk, 0)
    c = ref.RefChain("am", fs)
    ref_audio, ref_lens, _ = c.run(iq, blk)
    c.close()
    audio, lens, _, _ = restate.am_run(iq, fs, blk)
    assert list(lens) == list(ref_lens) and len(ref_audio) > 500
    # the resampler hands float32 samples to the decoder: a 1e-13 difference between r8brain's FFT convolution and the
    # restatement's direct one occasionally flips that rounding, which the float AGC chain carries into the audio
    assert np.abs(audio - ref_audio).max() <= 1e-7
    from airspy_fmradion_b200 import _capi
    L = _capi.lib()
    bl = np.full(nblk, blk, dtype=np.uint32)
    out = np.zeros(nblk, dtype=np.uint32)
    _capi.check(L.fmr_am_schedule(fs, 0, bl.ctypes.data, nblk, out.ctypes.data))
    assert list(out) == list(ref_lens)


def test_rates_without_tables_are_refused():
    """A ratio r8brain would design a chain for but whose tables are not shipped: FMR_ERR_UNSUPPORTED, never a guess."""
    from airspy_fmradion_b200 import _capi
    L = _capi.lib()
    bl = np.full(4, 2048, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    for fs in (2.0e6, 1.8e6, 250000.0):
        assert L.fmr_fm_schedule(fs, 1, 0, bl.ctypes.data, 4, None, out.ctypes.data) == 2  # FMR_ERR_UNSUPPORTED
    assert L.fmr_am_schedule(100000.0, 0, bl.ctypes.data, 4, out.ctypes.data) == 2
    for fs in RATES + [1.0e7]:
        assert L.fmr_fm_schedule(fs, 1, 0, bl.ctypes.data, 4, None, out.ctypes.data) == 0
