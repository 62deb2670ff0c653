"""48 kHz tail kernel (csrc/fmr_kernels.cuh k_fm_tail: DC block HighPassFilterIir Filter.cpp:304-311 + stereo matrix
FmDecode.cpp:194-220, one recurrence warp + helper warps per 32 channels) at its edges: channel counts that leave a
CTA partly filled, an output row pitch that rules out the 16-byte row stores, reference calls too short to produce
a 48 kHz sample (empty entries in the per-call flag list), mono and pilot_shift output modes. The same stream must come
out identically through every layout, and must match the oracle."""
import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import have_ref

pytestmark = pytest.mark.gpu
FS = 384000.0


def _run(iq, blocks_per_call, pad, **kw):
    """Device entry point, one call per entry of blocks_per_call (each a list of block lengths); `pad` is added to the
    audio row pitch (odd pad = rows that are not 16-byte aligned)."""
    import torch
    from airspy_fmradion_b200 import FmDecoder
    C = iq.shape[0]
    width = 2 if kw.get("stereo", True) else 1
    tmax = max(sum(b) for b in blocks_per_call)
    dec = FmDecoder(input_rate=FS, n_channels=C, max_samples_per_call=tmax, max_blocks_per_call=max(len(b) for b in blocks_per_call),
                    **kw)
    cap = int(tmax * 48000.0 / FS + 8) * width + pad
    d_iq = torch.from_numpy(iq).cuda()
    d_out = torch.full((C, cap), np.nan, dtype=torch.float64, device="cuda")
    sh = torch.cuda.current_stream().cuda_stream
    outs, lens, o = [], [], 0
    for bl in blocks_per_call:
        t = sum(bl)
        d_in = d_iq[:, o:o + t].contiguous()
        l = dec.process_device(d_in.data_ptr(), t, bl, d_out.data_ptr(), cap, sh)
        torch.cuda.synchronize()
        n = int(l.sum())
        outs.append(d_out[:, :n].cpu().numpy())
        tail = d_out[:, n:].cpu().numpy()
        assert np.isnan(tail).all(), "the kernel wrote past the samples of the call"
        d_out.fill_(np.nan)
        lens.append(l)
        o += t
    dec.close()
    return np.concatenate(outs, axis=1), np.concatenate(lens)


def _blocks():
    # ragged reference calls: 1-sample and 5-sample blocks give no 48 kHz output (empty calls), 8 samples give one
    rng = np.random.default_rng(7)
    calls = []
    for k in range(5):
        bl = []
        for _ in range(40):
            bl += [int(rng.choice([1, 5, 8, 64, 300, 1024, 2048]))]
        calls.append(bl)
    return calls


@pytest.mark.parametrize("kw", [dict(stereo=True), dict(stereo=True, pilot_shift=True), dict(stereo=False)],
                         ids=["stereo", "pilot_shift", "mono"])
def test_tail_layouts_agree_and_match_oracle(kw):
    calls = _blocks()
    n = sum(sum(b) for b in calls)
    base = [siggen.fm_stereo_iq(FS, n, c) for c in range(3)]
    C = 37  # one full CTA of 32 channels and one with 5 rows
    iq = np.stack([base[c % 3] for c in range(C)])
    a_even, l_even = _run(iq, calls, 0, **kw)
    a_odd, l_odd = _run(iq, calls, 1, **kw)
    assert list(l_even) == list(l_odd)
    assert np.array_equal(a_even, a_odd), "row pitch changes the audio"
    for c in range(3, C):
        assert np.array_equal(a_even[c], a_even[c % 3]), "channel %d differs from its twin" % c
    flat = [b for bl in calls for b in bl]
    # the oracle, block by block with the same partition
    if not have_ref():
        pytest.skip("the ragged partition needs the reference's block interface (oracle/_ref)")
    from oracle import ref
    ch = ref.RefChain("fm", FS, stereo=kw.get("stereo", True), pilot_shift=kw.get("pilot_shift", False))
    ref_audio, ref_lens, o = [], [], 0
    for b in flat:
        a = ch.process_block(iq[1][o:o + b])
        ref_audio.append(a)
        ref_lens.append(len(a))
        o += b
    ch.close()
    want = np.concatenate(ref_audio)
    assert list(l_even) == ref_lens
    d = a_even[1] - want
    print("tail %s: %d samples, max %.3e rms %.3e" % (kw, len(want), np.abs(d).max(), np.sqrt(np.mean(d * d))))
    assert np.abs(d).max() <= 2e-5 and np.sqrt(np.mean(d * d)) <= 5e-6
