"""GPU tests of the steps either side of the decoder path through the C ABI:
FileSource's sample formats decoded on the device, and the block loop's output stage (main.cpp:977-1002:
level metering, squelch gain, sink sample format) applied on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import fileio, siggen
from tests.oracle_select import oracle_fm_run

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair(cls, **kw):
    return cls(**kw), cls(**kw)


@pytest.mark.parametrize("fs,fmt", [(1.0e6, fileio.IQ_S16), (1.0e6, fileio.IQ_S8), (1.0e6, fileio.IQ_U8),
                                    (1.0e6, fileio.IQ_S24), (1.0e6, fileio.IQ_CF32), (384000.0, fileio.IQ_S16),
                                    (1.0e7, fileio.IQ_S24)])
def test_fm_device_decode_equals_sf_read_float(fs, fmt):
    """The file's bytes decoded on the device give bit-identical audio to the same samples converted on the host
    the way sf_read_float does (FileSource.cpp:491-531) and handed over as cf32."""
    from airspy_fmradion_b200 import FmDecoder
    blk, per, calls, nch = 2048, 48, 3, 3
    if fs > 5e6:
        per, calls = 160, 4  # first audio after 29 591 IF samples = 77 ms
    n = blk * per * calls
    raw = np.stack([fileio.quantize_iq(siggen.fm_stereo_iq(fs, n, c), fmt) for c in range(nch)])
    iq = np.stack([fileio.sf_read_float(raw[c], fmt) for c in range(nch)])
    a, b = _pair(FmDecoder, stereo=True, input_rate=fs, n_channels=nch, max_samples_per_call=blk * per,
                 max_blocks_per_call=per)
    esz = fileio.IQ_BYTES[fmt]
    total = 0
    for k in range(calls):
        s0, s1 = k * blk * per, (k + 1) * blk * per
        wa, la = a.process_blocks(iq[:, s0:s1], [blk] * per)
        wb, lb = b.process_blocks_io(raw[:, s0 * esz:s1 * esz], fmt, [blk] * per)
        assert list(la) == list(lb) and wb.dtype == np.float64
        assert np.array_equal(wa, wb)
        total += wa.shape[1]
    assert total > 1000
    assert a.stats(0).pll_lock_cnt == b.stats(0).pll_lock_cnt


def _block_if_rms(dec, lib, fs, start, bl, channel):
    """Per-block IF RMS from the decoder-input tap and the host schedule."""
    if_len = np.zeros(len(bl), dtype=np.uint32)
    au_len = np.zeros(len(bl), dtype=np.uint32)
    assert lib.fmr_fm_schedule(fs, 1, start, bl.ctypes.data, len(bl), if_len.ctypes.data, au_len.ctypes.data) == 0
    x = dec.tap_if(channel)
    assert len(x) == int(if_len.sum())
    out, o = [], 0
    for n in if_len:
        seg = x[o:o + n]
        o += n
        out.append(None if n == 0 else float(np.sqrt(np.mean(seg.real.astype(np.float64) ** 2 + seg.imag.astype(np.float64) ** 2))))
    return out, au_len


@pytest.mark.parametrize("out_fmt", [fileio.OUT_S16, fileio.OUT_F32, fileio.OUT_F64])
def test_fm_output_stage(out_fmt):
    """Levels, squelch gain and sink format on the device == main.cpp:977-1002 applied to the decoder's doubles.
    Channel 1 is 20 dB weaker and falls below the squelch level; the first blocks have no IF / no audio yet."""
    from airspy_fmradion_b200 import FmDecoder, _capi
    lib = _capi.lib()
    fs, blk, per, calls, nch = 1.0e6, 2048, 40, 3, 2
    n = blk * per * calls
    iq = np.stack([siggen.fm_stereo_iq(fs, n, 0), 0.1 * siggen.fm_stereo_iq(fs, n, 1)]).astype(np.complex64)
    a, b = _pair(FmDecoder, stereo=True, input_rate=fs, n_channels=nch, max_samples_per_call=blk * per,
                 max_blocks_per_call=per)
    squelch = 0.2
    bl = np.full(per, blk, dtype=np.uint32)
    seen_muted = seen_open = False
    for k in range(calls):
        seg = iq[:, k * blk * per:(k + 1) * blk * per]
        wa, la = a.process_blocks(seg, bl)
        wb, lb = b.process_blocks_io(seg.view(np.uint8).reshape(nch, -1), fileio.IQ_CF32, bl, out_format=out_fmt,
                                     squelch_level=squelch, gain=0.5)
        assert list(la) == list(lb)
        for c in range(nch):
            rms, au_len = _block_if_rms(a, lib, fs, k * blk * per, bl, c)
            assert list(au_len) == list(la)
            lv = b.block_levels(c)
            blocks, o = [], 0
            for m in la:
                blocks.append(wa[c, o:o + m])
                o += m
            # squelch decision from the device's own if_rms (the comparison is exact either way except at the threshold)
            dev_rms = [None if r is None else float(v) for r, v in zip(rms, lv[:, 0])]
            want, want_lv, _, _ = fileio.output_stage(blocks, dev_rms, out_fmt, squelch_level=squelch, gain=0.5)
            assert wb.dtype == want.dtype and np.array_equal(wb[c], want)
            for i, r in enumerate(rms):
                if r is None:
                    assert lv[i, 0] == -1.0
                    continue
                assert abs(lv[i, 0] - r) <= 2e-6 * max(r, 1e-3)
                assert lv[i, 3] == (0.5 if lv[i, 0] >= squelch else 0.0)
                if la[i]:
                    assert abs(lv[i, 1] - want_lv[i, 1]) <= 1e-6 and abs(lv[i, 2] - want_lv[i, 2]) <= 1e-6 * max(1.0, want_lv[i, 2])
            open_blocks = (lv[:, 3] > 0).sum()
            if c == 0:
                seen_open |= open_blocks > 0
                assert (lv[lv[:, 0] >= 0.3, 3] == 0.5).all()
            else:
                seen_muted |= (lv[lv[:, 0] >= 0, 3] == 0.0).all() and (lv[:, 0] >= 0).any()
    assert seen_open and seen_muted


def test_fm_io_capacity_and_argument_errors():
    from airspy_fmradion_b200 import FmDecoder, _capi
    d = FmDecoder(stereo=True, input_rate=1.0e6, n_channels=1, max_samples_per_call=4096, max_blocks_per_call=2)
    raw = np.zeros((1, 4096 * 4), dtype=np.uint8)
    lib = _capi.lib()
    bl = np.array([2048, 2048], dtype=np.uint32)
    out = np.zeros(4096, dtype=np.float64)
    lens = np.zeros(2, dtype=np.uint32)
    oc = _capi.OutputConfig(7, 0.0, 0.5)
    assert lib.fmr_fm_process_host_io(d._h, raw.ctypes.data, 9, 4096, bl.ctypes.data, 2, None, out.ctypes.data, 4096,
                                      lens.ctypes.data) == 1  # FMR_ERR_INVALID: unknown iq_format
    assert lib.fmr_fm_process_host_io(d._h, raw.ctypes.data, fileio.IQ_S16, 4096, bl.ctypes.data, 2, C.byref(oc),
                                      out.ctypes.data, 4096, lens.ctypes.data) == 1  # unknown out_format
    assert lib.fmr_fm_process_host_io(d._h, raw.ctypes.data, fileio.IQ_S16, 4095, bl.ctypes.data, 2, None,
                                      out.ctypes.data, 4096, lens.ctypes.data) == 1  # iq_stride < sum(block_len)
    d.process_blocks_io(raw, fileio.IQ_S16, [2048, 2048])
    with pytest.raises(_capi.FmrError):
        d.block_levels(0)  # the last call had no output stage
    out, lens = d.process_blocks_io(raw, fileio.IQ_S16, [2048, 2048], out_format=fileio.OUT_S16)
    lv = d.block_levels(0)
    assert lv.shape == (2, 4) and (lv[:, 0] <= 0).all()  # silence: if_rms 0, or -1 while the resampler fills


@pytest.mark.parametrize("mode,fs,fmt", [(2, 48000.0, fileio.IQ_S16), (2, 384000.0, fileio.IQ_S24), (1, 48000.0, fileio.IQ_U8)])
def test_am_nbfm_io(mode, fs, fmt):
    """48 kHz decoders: device-side sample decode bit-identical to host conversion; int16 sink == restatement."""
    from airspy_fmradion_b200 import AmDecoder
    blk, per, calls = 2048, 24, 3
    n = blk * per * calls
    x = siggen.am_iq(fs, n, 0) if mode == 2 else siggen.nbfm_iq(fs, n, 0)
    raw = fileio.quantize_iq(x, fmt)[None, :]
    iq = fileio.sf_read_float(raw[0], fmt)[None, :]
    kw = dict(mode=mode, input_rate=fs, n_channels=1, max_samples_per_call=blk * per, max_blocks_per_call=per)
    a, b, c = AmDecoder(**kw), AmDecoder(**kw), AmDecoder(**kw)
    esz = fileio.IQ_BYTES[fmt]
    bl = [blk] * per
    got_audio = 0
    for k in range(calls):
        s0, s1 = k * blk * per, (k + 1) * blk * per
        wa, la = a.process_blocks(iq[:, s0:s1], bl)
        wb, lb = b.process_blocks_io(raw[:, s0 * esz:s1 * esz], fmt, bl)
        assert list(la) == list(lb) and np.array_equal(wa, wb)
        wc, lc = c.process_blocks_io(raw[:, s0 * esz:s1 * esz], fmt, bl, out_format=fileio.OUT_S16, squelch_level=0.0,
                                     gain=0.5)
        lv = c.block_levels(0)
        assert list(lc) == list(la) and wc.dtype == np.int16
        assert np.array_equal(wc[0], fileio.sf_write_double(wa[0] * 0.5, fileio.OUT_S16))
        # the IF RMS of the last block is the decoder's own get_if_rms()
        last = [i for i in range(per) if lv[i, 0] >= 0]
        if last:
            assert abs(lv[last[-1], 0] - a.get_if_rms(0)) <= 2e-5 * max(a.get_if_rms(0), 1e-3)
        o = 0
        for i, m in enumerate(la):
            if m:
                f = wa[0, o:o + m].astype(np.float32)
                assert abs(lv[i, 2] - np.sqrt(np.mean(f.astype(np.float64) ** 2))) <= 2e-6
            o += m
        got_audio += wa.shape[1]
    assert got_audio > 1000


def test_cpp_file_to_file_drop_in(tmp_path):
    """C++: FileSource (16-bit WAV, 1 Msps) -> GPU (decode, FM stereo, levels, -6 dB, int16) -> SndfileOutput, against
    the oracle's audio put through the restated output stage: within one LSB of the 16-bit sink."""
    from scipy.io import wavfile
    exe = str(tmp_path / "file_decode")
    pkg = os.path.join(ROOT, "airspy_fmradion_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "file_decode.cpp"), "-o", exe,
                           "-L" + pkg, "-lfmradion_b200", "-Wl,-rpath," + pkg])
    fs, blk, nblk = 1.0e6, 2048, 150
    n = blk * nblk + 1000  # ragged last block
    raw = fileio.quantize_iq(siggen.fm_stereo_iq(fs, n, 0), fileio.IQ_S16)
    fin, fout = str(tmp_path / "iq.wav"), str(tmp_path / "audio.wav")
    fileio.write_wav(fin, raw, fileio.IQ_S16, int(fs))
    out = subprocess.run([exe, "filename=%s,blklen=%d" % (fin, blk), fout, "32", "-1"], capture_output=True, text=True)
    print(out.stdout.strip(), out.stderr.strip())
    assert out.returncode == 0
    rate, pcm = wavfile.read(fout)
    assert rate == 48000 and pcm.dtype == np.int16 and pcm.shape[1] == 2
    iq = fileio.sf_read_float(raw, fileio.IQ_S16)
    ref_audio, _ = oracle_fm_run(iq, fs, blk, stereo=True)
    want = fileio.sf_write_double(ref_audio * 0.5, fileio.OUT_S16)
    got = pcm.reshape(-1)
    assert len(got) == len(want) > 1000
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    print("file->file: %d values, max |diff| %d LSB, %d differ" % (len(got), d.max(), (d > 0).sum()))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    assert "muted=0" in out.stdout and "stereo=0" in out.stdout  # 0.31 s: the pilot PLL needs 0.5 s to report lock
    # squelch at -3 dB (IF RMS of this signal is ~ -7 dB): every block muted
    out = subprocess.run([exe, "filename=%s,blklen=%d" % (fin, blk), fout, "32", "3"], capture_output=True, text=True)
    assert out.returncode == 0
    rate, pcm = wavfile.read(fout)
    assert len(pcm.reshape(-1)) == len(want) and not pcm.any()
