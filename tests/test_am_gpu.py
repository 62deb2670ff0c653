"""GPU parity of the AM path (AmDecoder::process, ModType::AM) through the C ABI."""
import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_am_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fs,filt", [(384000.0, 0), (384000.0, 2), (48000.0, 3)])
def test_am_multichannel(fs, filt):
    from airspy_fmradion_b200 import AmDecoder
    blk, nblk, C = 2048, 150 if fs > 48000 else 40, 8
    iq = np.stack([siggen.am_iq(fs, blk * nblk, c) for c in range(C)])
    dec = AmDecoder(amfilter=filt, input_rate=fs, n_channels=C, max_samples_per_call=blk * 64)
    outs, lens = [], []
    for o in range(0, nblk, 64):
        k = min(64, nblk - o)
        a, l = dec.process_blocks(iq[:, o * blk:(o + k) * blk], [blk] * k)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    for c in (0, 3, C - 1):
        ref_audio, ref_lens = oracle_am_run(iq[c], fs, blk, filter=filt)
        assert list(lens) == list(ref_lens)
        d = audio[c] - ref_audio
        print("AM fs=%g filt=%d ch%d: n=%d max %.3e rms %.3e" % (fs, filt, c, len(d), np.abs(d).max(),
                                                               np.sqrt(np.mean(d * d))))
        assert np.abs(d).max() <= 2e-5 and np.sqrt(np.mean(d * d)) <= 5e-6
    s = dec.stats(0)
    assert s.decoder_calls == int((lens > 0).sum())
    assert 0.2 < s.af_agc_gain <= 1.5 and s.if_agc_gain > 1.0
