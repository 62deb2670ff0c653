"""GPU parity for the rate pairs added at the end of round 1 without GPU budget left (their kernels are the ones the
2.5 MHz and 1 MHz chains use; only the tables are new). Gated: run with FMR_EXPERIMENTAL_RATES=1, and once green flip
their `verified` flag in tools/gen_tables.py and drop the gate."""
import os

import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FMR_EXPERIMENTAL_RATES") != "1",
                                 reason="unverified rate pairs: set FMR_EXPERIMENTAL_RATES=1 (tools/next_round_ab.sh)")]


@pytest.mark.parametrize("fs", [3.0e6, 2.4e6, 2.048e6, 1.44e6, 1.2e6, 1152000.0, 960000.0, 912000.0, 768000.0])
def test_new_rate_matches_oracle(fs):
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
