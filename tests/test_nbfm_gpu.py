"""GPU parity of the NBFM path (NbfmDecoder::process, ModType::NBFM) through the C ABI:
against the oracle on seeded inputs, against the committed golden vectors, and edge cases."""
import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_nbfm_run
from tests.test_nbfm_cpu import CASES, GOLDEN, case_input, window_err

pytestmark = pytest.mark.gpu

TOL_MAX, TOL_RMS = 2e-5, 5e-6


@pytest.mark.parametrize("fs,filt,dev", [(384000.0, 0, 8000.0), (384000.0, 2, 8000.0), (48000.0, 3, 17000.0)])
def test_nbfm_multichannel(fs, filt, dev):
    from airspy_fmradion_b200 import NbfmDecoder
    blk, nblk, C = 2048, 150 if fs > 48000 else 40, 6
    iq = np.stack([siggen.nbfm_iq(fs, blk * nblk, c) for c in range(C)])
    dec = NbfmDecoder(nbfmfilter=filt, freq_dev=dev, input_rate=fs, n_channels=C, max_samples_per_call=blk * 64)
    outs, lens = [], []
    for o in range(0, nblk, 64):
        k = min(64, nblk - o)
        a, l = dec.process_blocks(iq[:, o * blk:(o + k) * blk], [blk] * k)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    for c in (0, 2, C - 1):
        ref_audio, ref_lens, st = oracle_nbfm_run(iq[c], fs, blk, filter=filt, freq_dev=dev)
        assert list(lens) == list(ref_lens)
        d = audio[c] - ref_audio
        print("NBFM fs=%g filt=%d ch%d: n=%d max %.3e rms %.3e" % (fs, filt, c, len(d), np.abs(d).max(),
                                                                 np.sqrt(np.mean(d * d))))
        assert len(d) > 5000 and np.abs(d).max() <= TOL_MAX and np.sqrt(np.mean(d * d)) <= TOL_RMS
        s = dec.stats(c)
        assert abs(s.if_rms - st.if_rms) < 1e-5 and abs(s.baseband_level - st.baseband_level) < 1e-5
        assert abs(s.tuning_offset - st.tuning_offset) < 0.05  # Hz
        assert abs(s.if_agc_gain - st.if_agc_gain) < 1e-4 * st.if_agc_gain
    assert dec.stats(0).decoder_calls == int((lens > 0).sum())


@pytest.mark.parametrize("name", sorted(CASES))
def test_nbfm_gpu_vs_golden(name):
    from airspy_fmradion_b200 import NbfmDecoder
    g = np.load(GOLDEN)
    fs, nblk, blk, skw, dkw = CASES[name]
    iq = case_input(name, g)[None, :]
    dec = NbfmDecoder(nbfmfilter=dkw.get("filter", 0), freq_dev=dkw.get("freq_dev", 8000.0), input_rate=fs,
                      fs4_shift=dkw.get("fs4", False), n_channels=1, max_samples_per_call=nblk * blk,
                      max_blocks_per_call=nblk)
    audio, lens = dec.process_blocks(iq, [blk] * nblk)
    assert list(lens) == list(g[name + "/lens"])
    err = window_err(g, name, audio[0])
    print(name, "max |gpu - golden| =", err)
    assert err <= TOL_MAX
    np.testing.assert_allclose(np.abs(audio[0]).sum(), g[name + "/audio_sum"][1], rtol=1e-5)
    s, st = g[name + "/stats"], dec.stats(0)
    assert abs(st.tuning_offset - s[0]) < 0.05 and abs(st.baseband_level - s[1]) < 1e-5
    assert abs(st.if_rms - s[2]) < 1e-5 and st.decoder_calls == int(s[4])


def test_nbfm_one_block_per_call_and_ragged():
    """Same stream as one super-block, as one block per call, and as ragged blocks (incl. empty ones):
    the per-call head-loop quirks of both FIR filters follow the partition, so the first two agree
    exactly and the ragged run matches the oracle driven with the same partition."""
    from airspy_fmradion_b200 import NbfmDecoder
    from oracle import ref
    fs, blk, nblk = 48000.0, 1024, 40
    iq = siggen.nbfm_iq(fs, blk * nblk, 4)[None, :]
    a = NbfmDecoder(input_rate=fs, n_channels=1, max_samples_per_call=blk * nblk, max_blocks_per_call=nblk)
    big, big_len = a.process_blocks(iq, [blk] * nblk)
    b = NbfmDecoder(input_rate=fs, n_channels=1, max_samples_per_call=blk, max_blocks_per_call=1)
    outs = [b.process_blocks(iq[:, i * blk:(i + 1) * blk], [blk])[0] for i in range(nblk)]
    small = np.concatenate(outs, axis=1)
    assert big.shape == small.shape and np.abs(big - small).max() <= 1e-12
    if not ref.available():
        pytest.skip("ragged partitions need the compiled reference")
    rng = np.random.default_rng(3)
    lens, left = [], blk * nblk
    while left > 0:
        n = min(int(rng.choice([0, 1, 5, 64, 300, 1024, 3000])), left)
        lens.append(n)
        left -= n
    c = NbfmDecoder(input_rate=fs, n_channels=1, max_samples_per_call=blk * nblk, max_blocks_per_call=len(lens))
    got, got_len = c.process_blocks(iq, lens)
    r = ref.RefChain("nbfm", fs)
    outs, o = [], 0
    for n in lens:
        outs.append(r.process_block(iq[0, o:o + n]) if n else np.empty(0))
        o += n
    want = np.concatenate(outs)
    assert list(got_len) == [len(x) for x in outs]
    assert np.abs(got[0] - want).max() <= TOL_MAX
