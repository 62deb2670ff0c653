"""Fused persistent front end (csrc/fmr_frontend.cuh: TMA-staged half-band producers -> FFT consumers in one kernel)
against the unfused kernels (k_hb_stream_tma / k_hb_cascade -> 1.25 MHz ring -> k_fdr): both run the same per-thread
arithmetic on the same absolute block grid, so the 384 kHz IF stream and the audio must be BIT-identical, for every
mix of call sizes; and against the oracle (reference IfResampler.cpp:37-79 + FmDecode.cpp:85-221)."""
import os

import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run

pytestmark = pytest.mark.gpu
FS, BLK = 1.0e7, 2048


def _run(calls, iq, env):
    from airspy_fmradion_b200 import FmDecoder
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        dec = FmDecoder(stereo=True, input_rate=FS, n_channels=iq.shape[0], max_samples_per_call=BLK * max(calls),
                        max_blocks_per_call=max(calls))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    dec.set_profiling(True)
    import torch
    # device entry point: one launch group per call (the host entry point cuts a call into copy/compute chunks that are
    # too short for whole blocks of the fused kernel)
    d_iq = torch.from_numpy(iq).cuda()
    cap = int(BLK * max(calls) * 48000.0 / FS) * 2 + 64
    d_out = torch.zeros((iq.shape[0], cap), dtype=torch.float64, device="cuda")
    sh = torch.cuda.current_stream().cuda_stream
    outs, lens, ifs, fused_calls, o = [], [], [[] for _ in range(iq.shape[0])], 0, 0
    for k in calls:
        d_in = d_iq[:, o * BLK:(o + k) * BLK].contiguous()
        l = dec.process_device(d_in.data_ptr(), k * BLK, [BLK] * k, d_out.data_ptr(), cap, sh)
        torch.cuda.synchronize()
        fused_calls += "if_frontend_fused" in dec.stage_times()
        outs.append(d_out[:, :int(l.sum())].cpu().numpy())
        lens.append(l)
        for c in range(iq.shape[0]):
            ifs[c].append(dec.tap_if(c))
        o += k
    return np.concatenate(outs, axis=1), np.concatenate(lens), [np.concatenate(x) for x in ifs], fused_calls


def test_fused_front_end_bit_identical_to_unfused_and_matches_oracle():
    calls = [150, 3, 1, 64, 2, 200, 1, 90, 33, 120]
    n = BLK * sum(calls)
    iq = np.stack([siggen.fm_stereo_iq(FS, n, c) for c in range(3)])
    a1, l1, if1, fused1 = _run(calls, iq, {})
    a0, l0, if0, fused0 = _run(calls, iq, {"FMR_FE": "0"})
    assert fused0 == 0 and fused1 >= 4, (fused0, fused1)  # the big calls take the fused kernel
    assert list(l1) == list(l0)
    for c in range(3):
        assert np.array_equal(if1[c].view(np.float32), if0[c].view(np.float32)), "IF stream differs on channel %d" % c
    assert np.array_equal(a1, a0)
    ref_audio, ref_lens, td, _ = oracle_fm_run(iq[1], FS, BLK, stereo=True, taps=("if",))
    assert list(l1) == list(ref_lens)
    want_if = np.concatenate(td["if"])
    e_if = np.abs(if1[1] - want_if).max()
    d = a1[1] - ref_audio
    print("fused front end: %d fused calls, IF max %.3e, audio max %.3e rms %.3e" % (fused1, e_if, np.abs(d).max(), np.sqrt(np.mean(d * d))))
    assert e_if < 2e-6 and np.abs(d).max() <= 2e-5 and np.sqrt(np.mean(d * d)) <= 5e-6


def test_fused_front_end_many_channels_one_call_per_step():
    """More channels than SMs (a CTA walks several channels, the TMA pipeline runs across channel boundaries) and the
    bench's call shape (329 blocks per call): fused == unfused, bit for bit, on every channel."""
    import torch
    nch, per, steps = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 5, 329, 2
    base = np.stack([siggen.fm_stereo_iq(FS, BLK * per * steps, c) for c in range(4)])
    iq = base[np.arange(nch) % 4] * (1.0 + 0.001 * (np.arange(nch) // 4))[:, None].astype(np.float32)
    iq = np.ascontiguousarray(iq.astype(np.complex64))
    a1, l1, _, fused1 = _run([per] * steps, iq, {})
    a0, l0, _, _ = _run([per] * steps, iq, {"FMR_FE": "0"})
    assert fused1 == steps and list(l1) == list(l0)
    assert np.array_equal(a1, a0)
