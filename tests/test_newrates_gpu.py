"""GPU parity for every rate pair the library serves beyond the BASELINE ones (reference: any `ifrate` the source
reports, main.cpp:673-729, IfResampler.cpp:25-35): same bar as the 10 MHz / 1 MHz chains. First green run on a B200:
profiles/pytest_newrates_r02.log."""
import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("fs", [6.0e6, 2.5e6])
def test_enabled_rate_matches_oracle(fs):
    """6 MHz (Airspy R2 alternative rate) and 2.5 MHz chains are served to every caller (`verified=1`): same parity bar
    as the 10 MHz / 1 MHz chains (main.cpp:673-729, IfResampler.cpp:25-35)."""
    _fm_rate_case(fs)


@pytest.mark.parametrize("fs", [3.0e6, 2.4e6, 2.048e6, 1.44e6, 1.2e6, 1152000.0, 960000.0, 912000.0, 768000.0])
def test_new_rate_matches_oracle(fs):
    _fm_rate_case(fs)


def _fm_rate_case(fs):
    from airspy_fmradion_b200 import FmDecoder
    blk, per = 2048, 64
    nblk = (int(np.ceil(0.75 * fs / blk)) + per - 1) // per * per  # past the PLL lock at 0.5 s
    iq = np.stack([siggen.fm_stereo_iq(fs, blk * nblk, c) for c in range(2)])
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=2, max_samples_per_call=blk * per, max_blocks_per_call=per)
    outs, lens = [], []
    for o in range(0, nblk, per):
        a, l = dec.process_blocks(iq[:, o * blk:(o + per) * blk], [blk] * per)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    for c in range(2):
        ref_audio, ref_lens = oracle_fm_run(iq[c], fs, blk, stereo=True)
        assert list(lens) == list(ref_lens)
        d = audio[c] - ref_audio
        print("fs=%g ch%d: n=%d max %.3e rms %.3e" % (fs, c, len(d), np.abs(d).max(), np.sqrt(np.mean(d * d))))
        assert np.abs(d).max() <= 2e-5 and np.sqrt(np.mean(d * d)) <= 5e-6
    assert dec.stereo_detected(0)


@pytest.mark.parametrize("fs", [192000.0, 256000.0, 768000.0, 1.0e6])
def test_new_am_rate_matches_oracle(fs):
    from airspy_fmradion_b200 import AmDecoder
    from tests.oracle_select import oracle_am_run
    blk, per = 2048, 32
    nblk = (int(np.ceil(0.5 * fs / blk)) + per - 1) // per * per
    iq = np.stack([siggen.am_iq(fs, blk * nblk, c) for c in range(2)])
    dec = AmDecoder(amfilter=0, input_rate=fs, n_channels=2, max_samples_per_call=blk * per, max_blocks_per_call=per)
    outs, lens = [], []
    for o in range(0, nblk, per):
        a, l = dec.process_blocks(iq[:, o * blk:(o + per) * blk], [blk] * per)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    for c in range(2):
        ref_audio, ref_lens = oracle_am_run(iq[c], fs, blk, filter=0)
        assert list(lens) == list(ref_lens)
        d = audio[c] - ref_audio
        print("AM fs=%g ch%d: n=%d max %.3e" % (fs, c, len(d), np.abs(d).max()))
        assert np.abs(d).max() <= 2e-5 and np.sqrt(np.mean(d * d)) <= 5e-6
