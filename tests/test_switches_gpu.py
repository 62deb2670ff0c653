"""Every environment switch the library still reads selects between two implementations of the same stage; each
non-default setting must decode the same stream as the default path (they differ by FP32 rounding in the IF stages and,
for FMR_AUDIO_FP64, in the audio filters) and as the oracle. One parametrised case per switch, so that a switch cannot rot.
Switches: fmr_host.cuh (Resampler::init), fmr_fm.cu (fmr_fm_create)."""
import os

import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run

pytestmark = pytest.mark.gpu
FS, BLK, CALLS = 1.0e7, 2048, [200, 40, 200, 150]  # 0.12 s: past both resamplers' start-up latency

SETTINGS = [
    {"FMR_FE": "0"},                                            # unfused front end
    {"FMR_FE_VARIANT": "1"},                                    # fused front end, split real / imaginary lanes
    {"FMR_FE_VARIANT": "2"},                                    # fused front end, eight consumer warps + setmaxnreg
    {"FMR_FE_MIN_BLOCKS": "4"},
    {"FMR_FE": "0", "FMR_FDR": "0"},                            # time-domain low-pass + polyphase bank (in-place FFT)
    {"FMR_FE": "0", "FMR_FDR": "0", "FMR_FFT_INPLACE": "0"},    # ... Stockham FFT
    {"FMR_FE": "0", "FMR_FDR": "0", "FMR_FFT_INPLACE": "0", "FMR_FFT_TW": "0"},
    {"FMR_FE": "0", "FMR_FDR": "0", "FMR_FUSE_FI": "0"},        # bank as its own launch
    {"FMR_FE": "0", "FMR_FDR": "0", "FMR_FFT": "0"},            # direct-form low-pass
    {"FMR_FE": "0", "FMR_HB_STREAM": "0"},                      # tiled half-band cascade
    {"FMR_FE": "0", "FMR_HBS_TMA": "0"},                        # streaming half-band cascade staged with cp.async
    {"FMR_AUDIO_FP64": "1"},                                    # audio half-bands, low-pass and pilot cut in FP64
    {"FMR_AUDIO_FP64": "1", "FMR_FFT_F64": "0"},                # ... with the audio low-pass as direct-form FIR
    {"FMR_CORE_FUSED": "0"},                                    # AGC / discriminator / PLL as separate launches
]


def _decode(iq, env):
    from airspy_fmradion_b200 import FmDecoder
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        dec = FmDecoder(stereo=True, input_rate=FS, n_channels=iq.shape[0], max_samples_per_call=BLK * max(CALLS),
                        max_blocks_per_call=max(CALLS))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    import torch
    d_iq = torch.from_numpy(iq).cuda()
    cap = int(BLK * max(CALLS) * 48000.0 / FS) * 2 + 64
    d_out = torch.zeros((iq.shape[0], cap), dtype=torch.float64, device="cuda")
    outs, lens, o = [], [], 0
    for k in CALLS:
        d_in = d_iq[:, o * BLK:(o + k) * BLK].contiguous()
        l = dec.process_device(d_in.data_ptr(), k * BLK, [BLK] * k, d_out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append(d_out[:, :int(l.sum())].cpu().numpy())
        lens.append(l)
        o += k
    return np.concatenate(outs, axis=1), np.concatenate(lens)


_cache = {}


def _baseline():
    if not _cache:
        iq = np.stack([siggen.fm_stereo_iq(FS, BLK * sum(CALLS), c) for c in range(2)])
        audio, lens = _decode(iq, {})
        ref_audio, ref_lens = oracle_fm_run(iq[1], FS, BLK, stereo=True)
        assert list(lens) == list(ref_lens) and len(ref_audio) > 1000
        assert np.abs(audio[1] - ref_audio).max() <= 2e-5
        _cache.update(iq=iq, audio=audio, lens=lens, ref=ref_audio)
    return _cache


@pytest.mark.parametrize("env", SETTINGS, ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_switch_decodes_the_same_stream(env):
    b = _baseline()
    audio, lens = _decode(b["iq"], env)
    assert list(lens) == list(b["lens"])
    d = np.abs(audio - b["audio"]).max()
    e = np.abs(audio[1] - b["ref"]).max()
    print("%s: max |switch - default| %.3e, max |switch - oracle| %.3e" % (env, d, e))
    assert d <= 5e-6 and e <= 2e-5


@pytest.mark.parametrize("mode", [2, 4, 6])  # AM, USB, CW (AmDecode.cpp:96-218)
def test_am_direct_form_channel_filter_decodes_the_same_stream(mode):
    """FMR_AM_FFT_FILTER=0: the 255- / 2049-tap channel filters in direct form instead of overlap-save FFT blocks + the
    head-loop correction (fmr_am.cu)."""
    from airspy_fmradion_b200 import AmDecoder
    fs, blk, nblk = 48000.0, 2048, 24
    iq = np.stack([siggen.am_iq(fs, blk * nblk, c) for c in range(2)])

    def run(env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            dec = AmDecoder(amfilter=0, mode=mode, input_rate=fs, n_channels=2, max_samples_per_call=blk * nblk,
                            max_blocks_per_call=nblk)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        return dec.process_blocks(iq, [blk] * nblk)

    a1, l1 = run({})
    a0, l0 = run({"FMR_AM_FFT_FILTER": "0"})
    assert list(l1) == list(l0) and a1.shape[1] > 1000
    d = np.abs(a1 - a0).max()
    print("mode %d: max |fft form - direct form| %.3e" % (mode, d))
    assert d <= 5e-6
