"""GPU path against the committed golden vectors (made from the compiled reference by
tools/gen_golden.py). These run even when oracle/_ref is absent on the GPU box."""
import numpy as np
import pytest

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(gu.CASES))
def test_gpu_vs_golden(name):
    from airspy_fmradion_b200 import AmDecoder, FmDecoder
    mode, fs, nblk, blk, skw, dkw = gu.CASES[name]
    iq = gu.case_input(name)[None, :]
    g = gu.golden()
    if mode == "fm":
        dec = FmDecoder(fmfilter=dkw.get("filter", 0), stereo=dkw.get("stereo", True),
                        multipath_stages=dkw.get("mpf_stages", 0), input_rate=fs, fs4_shift=dkw.get("fs4", False),
                        n_channels=1, max_samples_per_call=nblk * blk, max_blocks_per_call=nblk)
    else:
        dec = AmDecoder(input_rate=fs, n_channels=1, max_samples_per_call=nblk * blk, max_blocks_per_call=nblk)
    audio, lens = dec.process_blocks(iq, [blk] * nblk)
    assert list(lens) == list(g[name + "/lens"])
    tol = 1e-4 if dkw.get("mpf_stages") else 2e-5
    err = gu.check_window(name, "audio", audio[0], tol)
    print(name, "max |gpu - golden| =", err)
    assert err <= tol
    np.testing.assert_allclose(np.abs(audio[0]).sum(), g[name + "/audio_sum"][1], rtol=1e-5)
    if mode == "fm":
        s, st = g[name + "/stats"], dec.stats(0)
        assert st.stereo_detected == int(s[0]) and st.pll_lock_cnt == int(s[9]) and st.decoder_calls == int(s[10])
        assert abs(st.if_rms - s[4]) < 1e-5 and abs(st.baseband_level - s[2]) < 1e-5
        assert abs(st.tuning_offset - s[1]) < 0.05  # Hz
        if fs > 384000:
            assert gu.check_window(name, "if", dec.tap_if(0), 5e-6) <= 5e-6
        if dkw.get("mpf_stages"):
            ref_c = g[name + "/mpf_coeffs"]
            assert np.linalg.norm(dec.get_multipath_coefficients(0) - ref_c) <= 1e-3 * np.linalg.norm(ref_c)
