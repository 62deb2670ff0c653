"""CPU tests of the oracle itself: the plain-C restatement (oracle/restate) against the golden
vectors made from the compiled reference, and — when oracle/_ref was built here — against the
compiled reference live. This is what pins the oracle."""
import zlib

import numpy as np
import pytest

from oracle import ref, restate
from tests import golden_util as gu


@pytest.mark.parametrize("name", ["fm_mono_1M", "fm_stereo_10M", "fm_stereo_384k_fs4_narrow", "fm_stereo_384k_E8",
                                  "fm_stereo_384k_blk777", "fm_stereo_1M"])
def test_restatement_fm_vs_golden(name):
    mode, fs, nblk, blk, skw, dkw = gu.CASES[name]
    iq = gu.case_input(name)
    audio, lens, td, st = restate.fm_run(iq, fs, blk, taps=("if",) if fs > 384000 else (), **dkw)
    g = gu.golden()
    assert list(lens) == list(g[name + "/lens"])
    # double-precision path: 1e-9; float FIR / adaptive filter stages are sensitive to the
    # compiler's float summation order: 2e-8 / 1e-6
    tol = 1e-6 if dkw.get("mpf_stages") else (2e-8 if dkw.get("filter") else 1e-9)
    assert gu.check_window(name, "audio", audio, tol) <= tol
    np.testing.assert_allclose([audio.sum(), np.abs(audio).sum()], g[name + "/audio_sum"], rtol=1e-6, atol=1e-6)
    if fs > 384000:
        ifs = np.concatenate(td["if"])
        assert gu.check_window(name, "if", ifs, 1e-6) <= 1e-6
    s = g[name + "/stats"]
    assert st.stereo_detected == int(s[0])
    assert abs(st.baseband_level - s[2]) < 1e-6 and abs(st.pilot_level - s[3]) < 1e-7
    assert abs(st.if_rms - s[4]) < 1e-6 and abs(st.agc_gain - s[6]) < 1e-4 * s[6]
    assert st.pll_lock_cnt == int(s[9]) and st.decoder_calls == int(s[10])
    if dkw.get("mpf_stages"):
        ref_c = g[name + "/mpf_coeffs"]
        assert np.linalg.norm(st.mpf_coeffs - ref_c) <= 1e-4 * np.linalg.norm(ref_c)


def test_restatement_am_vs_golden():
    name = "am_384k"
    mode, fs, nblk, blk, skw, dkw = gu.CASES[name]
    iq = gu.case_input(name)
    audio, lens, _, st = restate.am_run(iq, fs, blk)
    g = gu.golden()
    assert list(lens) == list(g[name + "/lens"])
    assert gu.check_window(name, "audio", audio, 1e-7) <= 1e-7
    s = g[name + "/stats"]
    assert abs(st.baseband_level - s[0]) < 1e-6 and abs(st.af_agc_gain - s[1]) < 1e-6
    assert abs(st.if_agc_gain - s[2]) < 1e-4 * s[2] and abs(st.if_rms - s[3]) < 1e-6


@pytest.mark.parametrize("chain", [(1e7, 384000.0, 0), (6e6, 384000.0, 0), (2.5e6, 384000.0, 0), (1e6, 384000.0, 0),
                                   (384000.0, 48000.0, 0), (384000.0, 48000.0, 1)])
def test_restatement_resampler_vs_golden(chain):
    src, dst, kind = chain
    g = gu.golden()
    key = "r8b_%d_%d_%d" % (src, dst, kind)
    want, lens = g[key + "/out"], g[key + "/lens"]
    n = {1e7: 150000, 6e6: 120000, 2.5e6: 60000, 1e6: 40000}.get(src, 50000)
    x = np.random.Generator(np.random.PCG64(99)).standard_normal(n)
    if zlib.crc32(x.tobytes()) != int(g[key + "/crc"][0]):
        pytest.skip("numpy generator stream differs from the one the golden vectors were made with")
    r = restate.R8(src, dst, kind)
    ys = [r.process(x[o:o + 3000]) for o in range(0, n, 3000)]
    assert [len(y) for y in ys] == list(lens)
    got = np.concatenate(ys)
    assert np.abs(got - want).max() < 1e-9
    # the closed-form release schedule agrees with the streamed counts
    cum, tot = 0, 0
    for o, l in zip(range(0, n, 3000), lens):
        cum += min(3000, n - o)
        tot += int(l)
        assert restate.chain_out(src, dst, kind, cum) == tot


def test_fast_atan2f_known_answers():
    """fast_atan2f restatement against the reference's values on a grid (bit exact)."""
    import ctypes as C
    g = gu.golden()
    grid, val = g["fast_atan2f/grid"], g["fast_atan2f/val"]
    L = restate.lib()
    # the C function is static; exercise it through the PLL would be indirect, so check the table
    # restatement here: recompute with the same algorithm in numpy float32
    from tests.np_fast_atan2 import fast_atan2f_np
    got = np.array([[fast_atan2f_np(np.float32(y), np.float32(x)) for x in grid] for y in grid], dtype=np.float32)
    assert np.array_equal(got, val)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libfmref.so not built in this environment")
def test_restatement_vs_compiled_reference_live():
    from oracle import siggen
    fs, blk = 2.5e6, 4096
    iq = siggen.fm_stereo_iq(fs, blk * 200, 7)
    a, la, _, sa = restate.fm_run(iq, fs, blk, stereo=True, deemphasis_us=75.0)
    c = ref.RefChain("fm", fs, stereo=True, deemphasis_us=75.0)
    b, lb, _ = c.run(iq, blk)
    assert list(la) == list(lb)
    assert np.abs(a - b).max() < 1e-9
    sb = c.stats()
    assert abs(sa.if_rms - sb.if_rms) < 1e-6 and sa.pll_lock_cnt == sb.pll_lock_cnt
