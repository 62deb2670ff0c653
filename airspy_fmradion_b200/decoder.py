"""Host-side mirror of the reference's decoder block API over the C ABI.

`FmDecoder` / `AmDecoder` keep the reference's names, constructor argument meaning and
getters (include/FmDecode.h:34-105, include/AmDecode.h:32-65) so that tests read like the
reference's usage in main.cpp:812-830,953-974. Differences, all additive:
  * a decoder holds `n_channels` independent streams decoded in lock step;
  * the front end the reference runs just before the decoder (FourthConverterIQ, IfResampler;
    main.cpp:912-926) is part of the object (`input_rate`, `fs4_shift`), because the GPU
    path absorbs it;
  * `process_blocks` hands over many source blocks at once together with the block
    partition; `process` is the one-block form with the reference's signature.
No torch types here; device-pointer entry points take integer addresses.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check


class _Base:
    _h = None

    def close(self):
        if self._h:
            self._destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shared process plumbing -------------------------------------------------------
    def _blocks(self, block_len):
        bl = np.ascontiguousarray(block_len, dtype=np.uint32)
        return bl, int(bl.sum())

    def query_output(self, block_len):
        bl, _ = self._blocks(block_len)
        tot = C.c_uint64(0)
        lens = np.zeros(len(bl), dtype=np.uint32)
        check(self._query(self._h, bl.ctypes.data, len(bl), C.byref(tot), lens.ctypes.data))
        return int(tot.value), lens

    def _out_buffer(self, out, out_total):
        if out is None:
            return np.zeros((self.n_channels, max(out_total, 1)), dtype=np.float64)
        assert out.dtype == np.float64 and out.ndim == 2 and out.shape[0] == self.n_channels
        assert out.shape[1] >= out_total and out.flags["C_CONTIGUOUS"]
        return out

    def process_blocks(self, iq, block_len, out=None):
        """iq: complex64 [C, T] (or [T] for one channel); returns (audio [C, n], audio_len[n_blocks]).
        `out`: optional caller-owned float64 [C, >=n] buffer (e.g. pinned memory) to write into."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq[None, :]
        assert iq.shape[0] == self.n_channels
        bl, total = self._blocks(block_len)
        assert total <= iq.shape[1]
        out_total, _ = self.query_output(bl)
        audio = self._out_buffer(out, out_total)
        lens = np.zeros(len(bl), dtype=np.uint32)
        check(self._process_host(self._h, iq.ctypes.data, iq.shape[1], bl.ctypes.data, len(bl),
                                 audio.ctypes.data, audio.shape[1], lens.ctypes.data))
        return audio[:, :out_total], lens

    def process_device(self, d_iq_ptr, iq_stride, block_len, d_audio_ptr, audio_stride, stream=0):
        """Device-pointer form (addresses as ints, e.g. torch.Tensor.data_ptr()); asynchronous on
        `stream` (a cudaStream_t handle as int). Returns audio_len per block."""
        bl, _ = self._blocks(block_len)
        lens = np.zeros(len(bl), dtype=np.uint32)
        check(self._process_device(self._h, d_iq_ptr, iq_stride, bl.ctypes.data, len(bl), d_audio_ptr,
                                   audio_stride, lens.ctypes.data, stream))
        return lens

    # ---- FileSource sample formats in, the block loop's output stage out -------
    _OUT_DTYPES = {_capi.OUT_F64: np.float64, _capi.OUT_F32: np.float32, _capi.OUT_S16: np.int16}

    def process_blocks_io(self, raw, iq_format, block_len, out_format=None, squelch_level=0.0, gain=0.5, out=None):
        """raw: the file's own sample bytes per channel, uint8 [C, T * bytes_per_complex_sample] (or any
        C-contiguous array with that many bytes per row: int16 [C, T, 2], complex64 [C, T], ...).
        out_format None: the decoder's doubles unchanged; otherwise main.cpp:977-1002 is applied on the
        device (levels, squelch gain, sink sample format) and `block_levels()` holds the per-block levels.
        Returns (audio [C, n] in the sink dtype, audio_len[n_blocks])."""
        raw = np.ascontiguousarray(raw)
        if raw.ndim == 1 or (self.n_channels == 1 and raw.shape[0] != 1):
            raw = raw[None, ...]
        assert raw.shape[0] == self.n_channels
        esz = _capi.IQ_BYTES[iq_format]
        row_bytes = raw.nbytes // self.n_channels
        assert row_bytes % esz == 0
        stride = row_bytes // esz
        bl, total = self._blocks(block_len)
        assert total <= stride
        out_total, _ = self.query_output(bl)
        dt = self._OUT_DTYPES[_capi.OUT_F64 if out_format is None else out_format]
        if out is None:
            out = np.zeros((self.n_channels, max(out_total, 1)), dtype=dt)
        assert out.dtype == dt and out.ndim == 2 and out.shape[0] == self.n_channels and out.shape[1] >= out_total
        oc = None if out_format is None else C.byref(_capi.OutputConfig(int(out_format), float(squelch_level), float(gain)))
        lens = np.zeros(len(bl), dtype=np.uint32)
        check(self._process_host_io(self._h, raw.ctypes.data, int(iq_format), stride, bl.ctypes.data, len(bl), oc,
                                    out.ctypes.data, out.shape[1], lens.ctypes.data))
        self._last_io_blocks = len(bl)
        return out[:, :out_total], lens

    def process_device_io(self, d_raw_ptr, iq_format, iq_stride, block_len, d_audio_ptr, audio_stride, out_format=None,
                          squelch_level=0.0, gain=0.5, stream=0):
        """Device-pointer form of process_blocks_io (addresses as ints); asynchronous on `stream`. iq_stride in complex
        samples, audio_stride in values of the sink format. Returns audio_len per block."""
        bl, _ = self._blocks(block_len)
        oc = None if out_format is None else C.byref(_capi.OutputConfig(int(out_format), float(squelch_level), float(gain)))
        lens = np.zeros(len(bl), dtype=np.uint32)
        check(self._process_device_io(self._h, d_raw_ptr, int(iq_format), iq_stride, bl.ctypes.data, len(bl), oc,
                                      d_audio_ptr, audio_stride, lens.ctypes.data, stream))
        self._last_io_blocks = len(bl)
        return lens

    def block_levels(self, channel=0):
        """Per-block (if_rms, audio_mean, audio_rms, gain) of the last process_blocks_io call with an out_format."""
        n = self._last_io_blocks
        arr = (_capi.BlockLevel * n)()
        check(self._block_levels(self._h, channel, arr, n))
        return np.array([(a.if_rms, a.audio_mean, a.audio_rms, a.gain) for a in arr], dtype=np.float32).reshape(n, 4)

    def set_profiling(self, enable=True):
        check(self._set_profiling(self._h, int(enable)))

    def stage_times(self):
        """{stage name: ms} of the last process call (needs set_profiling(True))."""
        ms = (C.c_float * 16)()
        names = (C.c_char_p * 16)()
        n = C.c_uint32(0)
        check(self._stage_times(self._h, ms, names, 16, C.byref(n)))
        return {names[i].decode(): float(ms[i]) for i in range(n.value)}

    def process(self, samples_in):
        """Reference signature: one block in, audio out (FmDecode.h:74 / AmDecode.h:55)."""
        samples_in = np.asarray(samples_in)
        n = samples_in.shape[-1]
        audio, _ = self.process_blocks(samples_in, [n])
        return audio[0] if samples_in.ndim == 1 else audio


class FmDecoder(_Base):
    # Static constants (include/FmDecode.h:38-47)
    sample_rate_if = 384000.0
    sample_rate_pcm = 48000.0
    freq_dev = 75000.0
    bandwidth_pcm = 15000.0
    pilot_freq = 19000.0
    deemphasis_time_eu = 50.0
    deemphasis_time_na = 75.0

    def __init__(self, fmfilter=0, stereo=True, deemphasis=50.0, pilot_shift=False, multipath_stages=0, *,
                 input_rate=384000.0, fs4_shift=False, n_channels=1, max_samples_per_call=1 << 20,
                 max_blocks_per_call=4096, device=0, fmfilter_coeff=None):
        """fmfilter: 0 none (FilterType Default/Wide), 1 medium, 2 narrow (main.cpp:785-810), or pass
        fmfilter_coeff (the reference's `fmfilter_coeff` vector); the other arguments are FmDecoder's
        (FmDecode.h:49-64). Environment at creation: FMR_AUDIO_FP64=1 runs the audio resamplers and the pilot-cut FIR
        in double (default float; the recurrences are double either way), see include/fmradion_b200.h."""
        L = _capi.lib()
        self._destroy, self._query = L.fmr_fm_destroy, L.fmr_fm_query_output
        self._process_host, self._process_device = L.fmr_fm_process_host, L.fmr_fm_process_device
        self._process_host_io, self._block_levels = L.fmr_fm_process_host_io, L.fmr_fm_block_levels
        self._process_device_io = L.fmr_fm_process_device_io
        self._set_profiling, self._stage_times = L.fmr_fm_set_profiling, L.fmr_fm_stage_times
        self.n_channels = int(n_channels)
        self.stereo = bool(stereo)
        self.multipath_stages = int(multipath_stages)
        coeff = None
        if fmfilter_coeff is not None:
            coeff = np.ascontiguousarray(fmfilter_coeff, dtype=np.float32)
            fmfilter = 3
        cfg = _capi.FmConfig(float(input_rate), int(fs4_shift), int(fmfilter), int(stereo), float(deemphasis),
                             int(pilot_shift), int(multipath_stages), int(n_channels),
                             int(max_samples_per_call), int(max_blocks_per_call), int(device),
                             coeff.ctypes.data if coeff is not None else None, len(coeff) if coeff is not None else 0)
        h = C.c_void_p()
        check(L.fmr_fm_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def process_blocks_i16(self, iq_i16, block_len, out=None):
        """iq_i16: int16 [C, T, 2] (re,im) pairs as FileSource reads them from a 16-bit WAV."""
        iq = np.ascontiguousarray(iq_i16, dtype=np.int16)
        assert iq.ndim == 3 and iq.shape[0] == self.n_channels and iq.shape[2] == 2
        bl, total = self._blocks(block_len)
        assert total <= iq.shape[1]
        out_total, _ = self.query_output(bl)
        audio = self._out_buffer(out, out_total)
        lens = np.zeros(len(bl), dtype=np.uint32)
        check(_capi.lib().fmr_fm_process_host_i16(self._h, iq.ctypes.data, iq.shape[1], bl.ctypes.data, len(bl),
                                                  audio.ctypes.data, audio.shape[1], lens.ctypes.data))
        return audio[:, :out_total], lens

    def stats(self, channel=0):
        s = _capi.FmStats()
        check(_capi.lib().fmr_fm_stats(self._h, channel, C.byref(s)))
        return s

    def stereo_detected(self, channel=0):
        return bool(self.stats(channel).stereo_detected)

    def get_tuning_offset(self, channel=0):
        return self.stats(channel).tuning_offset

    def get_baseband_level(self, channel=0):
        return self.stats(channel).baseband_level

    def get_pilot_level(self, channel=0):
        return self.stats(channel).pilot_level

    def get_if_rms(self, channel=0):
        return self.stats(channel).if_rms

    def get_multipath_error(self, channel=0):
        return self.stats(channel).multipath_error

    def get_pps_events(self, channel=0):
        ev = (_capi.PpsEvent * 16)()
        n = C.c_uint32(0)
        check(_capi.lib().fmr_fm_pps_events(self._h, channel, ev, 16, C.byref(n)))
        return [(e.pps_index, e.sample_index, e.block_position, e.block) for e in ev[:n.value]]

    def get_multipath_coefficients(self, channel=0):
        n = 4 * self.multipath_stages + 1
        buf = np.zeros(2 * n, dtype=np.float32)
        check(_capi.lib().fmr_fm_coeffs(self._h, channel, buf.ctypes.data, n))
        return buf.view(np.complex64)

    def block_flags(self, n_blocks, channel=0):
        f = np.zeros(n_blocks, dtype=np.uint8)
        check(_capi.lib().fmr_fm_block_flags(self._h, channel, f.ctypes.data, n_blocks))
        return f

    def tap_if(self, channel=0):
        n = C.c_uint64(0)
        check(_capi.lib().fmr_fm_tap_if(self._h, channel, None, 0, C.byref(n)))
        buf = np.zeros(2 * max(int(n.value), 1), dtype=np.float32)
        check(_capi.lib().fmr_fm_tap_if(self._h, channel, buf.ctypes.data, int(n.value), C.byref(n)))
        return buf[:2 * int(n.value)].view(np.complex64)

    def last_launches(self):
        return int(_capi.lib().fmr_fm_last_launches(self._h))

    def describe(self):
        """Which implementation of every stage this handle selected (fmr_fm_describe)."""
        buf = C.create_string_buffer(512)
        _capi.lib().fmr_fm_describe(self._h, buf, 512)
        return buf.value.decode()

    def last_plan(self):
        """How the last call's IF front end ran, per channel (fmr_fm_last_plan)."""
        p = (C.c_uint64 * 5)()
        check(_capi.lib().fmr_fm_last_plan(self._h, p))
        return {"fused_blocks": int(p[0]), "unfused_blocks": int(p[1]), "unfused_halfband_outputs": int(p[2]),
                "block_in": int(p[3]), "block_out": int(p[4])}


class AmDecoder(_Base):
    # Static constants (include/AmDecode.h:35-40)
    sample_rate_pcm = 48000.0
    internal_rate_pcm = 48000.0
    bandwidth_pcm = 4500.0
    deemphasis_time = 100.0
    MODTYPE_AM = 2  # include/SoftFM.h:56

    def __init__(self, amfilter=0, mode=2, *, input_rate=48000.0, fs4_shift=False, n_channels=1,
                 max_samples_per_call=1 << 20, max_blocks_per_call=4096, device=0, amfilter_coeff=None):
        """amfilter: 0 default, 1 medium, 2 narrow, 3 wide (main.cpp:785-810); mode: ModType value
        (include/SoftFM.h:49): 2 AM, 3 DSB, 4 USB, 5 LSB, 6 CW, 7 WSPR."""
        L = _capi.lib()
        self._destroy, self._query = L.fmr_am_destroy, L.fmr_am_query_output
        self._process_host, self._process_device = L.fmr_am_process_host, L.fmr_am_process_device
        self._process_host_io, self._block_levels = L.fmr_am_process_host_io, L.fmr_am_block_levels
        self._process_device_io = L.fmr_am_process_device_io
        self._set_profiling, self._stage_times = L.fmr_am_set_profiling, L.fmr_am_stage_times
        self.n_channels = int(n_channels)
        coeff = None
        if amfilter_coeff is not None:
            coeff = np.ascontiguousarray(amfilter_coeff, dtype=np.float32)
            amfilter = 4
        cfg = _capi.AmConfig(float(input_rate), int(fs4_shift), int(amfilter), int(mode), int(n_channels),
                             int(max_samples_per_call), int(max_blocks_per_call), int(device),
                             coeff.ctypes.data if coeff is not None else None, len(coeff) if coeff is not None else 0,
                             float(getattr(self, "_freq_dev", 0.0)))
        h = C.c_void_p()
        check(L.fmr_am_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def stats(self, channel=0):
        s = _capi.AmStats()
        check(_capi.lib().fmr_am_stats(self._h, channel, C.byref(s)))
        return s

    def get_baseband_level(self, channel=0):
        return self.stats(channel).baseband_level

    def get_af_agc_current_gain(self, channel=0):
        return self.stats(channel).af_agc_gain

    def get_if_agc_current_gain(self, channel=0):
        return self.stats(channel).if_agc_gain

    def get_if_rms(self, channel=0):
        return self.stats(channel).if_rms

    def last_launches(self):
        return int(_capi.lib().fmr_am_last_launches(self._h))


class NbfmDecoder(AmDecoder):
    """Mirror of the reference's NbfmDecoder (include/NbfmDecode.h:30-63): narrow-band FM at 48 kHz,
    mono audio. Same handle type as AmDecoder on the C side (mode = ModType::NBFM)."""
    sample_rate_pcm = 48000.0
    internal_rate_pcm = 48000.0
    freq_dev_normal = 8000.0   # include/NbfmDecode.h:38
    freq_dev_wide = 17000.0    # include/NbfmDecode.h:41
    MODTYPE_NBFM = 1           # include/SoftFM.h:49

    def __init__(self, nbfmfilter=0, freq_dev=8000.0, *, input_rate=48000.0, fs4_shift=False, n_channels=1,
                 max_samples_per_call=1 << 20, max_blocks_per_call=4096, device=0, nbfmfilter_coeff=None):
        """nbfmfilter: 0 default, 1 medium, 2 narrow, 3 wide (main.cpp:785-810); freq_dev: full-scale deviation, Hz."""
        self._freq_dev = float(freq_dev)
        super().__init__(nbfmfilter, self.MODTYPE_NBFM, input_rate=input_rate, fs4_shift=fs4_shift,
                         n_channels=n_channels, max_samples_per_call=max_samples_per_call,
                         max_blocks_per_call=max_blocks_per_call, device=device, amfilter_coeff=nbfmfilter_coeff)

    def get_tuning_offset(self, channel=0):
        return self.stats(channel).tuning_offset
