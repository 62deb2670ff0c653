"""TEST INFRASTRUCTURE ONLY (never imported by the product path) — numpy restatement of the two steps either
side of the decoder path.

  * what `FileSource::get_sf_read_float` hands to the block loop (sfmbase/FileSource.cpp:491-531): libsndfile's
    `sf_read_float` with default normalisation, for the sub-types FileSource accepts (FileSource.cpp:206-216).
    libsndfile is an un-vendored, unpinned dependency that is absent from /root/reference and from this image;
    the conversions restated here are the published ones of its pcm.c (sc2f_array, uc2f_array, les2f_array,
    let2f_array: integer value times 2^-(bits-1), PCM_U8 re-centred by 128) and float32.c (FLOAT unchanged).
    Parity at this boundary is anchored on FileSource's own call site and independently on Python's `wave` /
    scipy.io.wavfile readers of the same files (tests/test_io_cpu.py) — "parity unpinned" against libsndfile itself.
  * the block loop's output stage, main.cpp:977-1002: `if_level`, audio level via `Utility::samples_mean_rms`
    (Utility.h:135-152) on the float copy, `Utility::adjust_gain` (Utility.h:307-312) with the IF squelch, and
    `SndfileOutput::write` = `sf_write_double` (AudioOutput.cpp:153-167): PCM_16 = lrint(x * 32767) without clipping,
    FLOAT = (float)x (libsndfile defaults norm_double on, add_clipping off).

Also writes the container fixtures (WAV, WAVEX, W64, RAW) the tests feed to the C++ FileSource.
"""
import struct

import numpy as np

IQ_CF32, IQ_S16, IQ_S8, IQ_U8, IQ_S24 = range(5)
OUT_F64, OUT_F32, OUT_S16 = range(3)
IQ_BYTES = {IQ_CF32: 8, IQ_S16: 4, IQ_S8: 2, IQ_U8: 2, IQ_S24: 6}


def quantize_iq(iq, fmt, scale=0.9):
    """complex64 [T] -> the file's own bytes (uint8 [T * bytes]) for sample format `fmt` (test signal generator)."""
    x = np.empty(2 * len(iq), dtype=np.float64)
    x[0::2], x[1::2] = iq.real, iq.imag
    x = np.clip(x * scale, -1.0, 1.0)
    if fmt == IQ_CF32:
        return x.astype("<f4").view(np.uint8)
    if fmt == IQ_S16:
        return np.clip(np.rint(x * 32767), -32768, 32767).astype("<i2").view(np.uint8)
    if fmt == IQ_S8:
        return np.clip(np.rint(x * 127), -128, 127).astype(np.int8).view(np.uint8)
    if fmt == IQ_U8:
        return (np.clip(np.rint(x * 127), -128, 127) + 128).astype(np.uint8)
    if fmt == IQ_S24:
        v = np.clip(np.rint(x * 8388607), -8388608, 8388607).astype("<i4")
        return np.ascontiguousarray(v.view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1)
    raise ValueError(fmt)


def sf_read_float(raw, fmt):
    """The file's bytes -> complex64, as sf_read_float + FileSource.cpp:518-528 deliver them."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
    if fmt == IQ_CF32:
        f = raw.view("<f4").astype(np.float32)
    elif fmt == IQ_S16:
        f = raw.view("<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    elif fmt == IQ_S8:
        f = raw.view(np.int8).astype(np.float32) * np.float32(1.0 / 128.0)
    elif fmt == IQ_U8:
        f = (raw.astype(np.int32) - 128).astype(np.float32) * np.float32(1.0 / 128.0)
    elif fmt == IQ_S24:
        b = raw.reshape(-1, 3).astype(np.uint32)
        v = ((b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)).astype(np.uint32).view(np.int32)
        f = v.astype(np.float32) * np.float32(1.0 / 2147483648.0)
    else:
        raise ValueError(fmt)
    n = len(f) // 2
    return (f[0:2 * n:2] + 1j * f[1:2 * n:2]).astype(np.complex64)


def sf_write_double(x, out_fmt):
    x = np.asarray(x, dtype=np.float64)
    if out_fmt == OUT_F64:
        return x.copy()
    if out_fmt == OUT_F32:
        return x.astype(np.float32)
    if out_fmt == OUT_S16:  # lrint, then the implicit long -> short truncation
        return np.rint(x * 32767.0).astype(np.int64).astype(np.int16)
    raise ValueError(out_fmt)


def output_stage(audio_blocks, if_rms_blocks, out_fmt, squelch_level=0.0, gain=0.5):
    """main.cpp:977-1002 over a list of per-block audio arrays (float64) and per-block IF RMS values (None where the
    block produced no IF samples). Returns (sink samples concatenated, levels [n_blocks, 4], if_level, audio_level)."""
    if_level = np.float32(0)
    audio_level = np.float32(0)
    out, levels = [], []
    for a, r in zip(audio_blocks, if_rms_blocks):
        if r is None:
            levels.append((-1.0, 0.0, 0.0, 0.0))
            continue
        if_level = np.float32(0.75 * float(if_level) + 0.25 * float(r))
        g = gain if float(r) >= squelch_level else 0.0
        if len(a) == 0:
            levels.append((r, 0.0, 0.0, g))
            continue
        f = a.astype(np.float32)
        mean = np.float32(f.sum(dtype=np.float32) / np.float32(len(f)))
        rms = np.float32(np.sqrt(np.float32(np.dot(f, f)) / np.float32(len(f))))
        audio_level = np.float32(0.95 * float(audio_level) + 0.05 * float(rms))
        out.append(sf_write_double(a * g, out_fmt))
        levels.append((r, mean, rms, g))
    dt = {OUT_F64: np.float64, OUT_F32: np.float32, OUT_S16: np.int16}[out_fmt]
    cat = np.concatenate(out) if out else np.zeros(0, dtype=dt)
    return cat, np.array(levels, dtype=np.float32).reshape(-1, 4), float(if_level), float(audio_level)


# ---------------------------------------------------------------- container fixtures

def _fmt_body(tag, channels, rate, bits, extensible=False):
    align = channels * bits // 8
    if not extensible:
        return struct.pack("<HHIIHH", tag, channels, rate, rate * align, align, bits)
    guid_tail = bytes([0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71])
    return (struct.pack("<HHIIHH", 0xFFFE, channels, rate, rate * align, align, bits) +
            struct.pack("<HHI", 22, bits, 3) + struct.pack("<H", tag) + guid_tail)


def _tag_bits(fmt):
    return {IQ_CF32: (3, 32), IQ_S16: (1, 16), IQ_U8: (1, 8), IQ_S24: (1, 24)}[fmt]


def write_wav(path, raw, fmt, rate, extensible=False, junk=False, data_len_override=None):
    """RIFF/WAVE with 2 channels (I, Q). junk: put a LIST chunk of odd length before 'fmt ' (padding rule)."""
    tag, bits = _tag_bits(fmt)
    raw = bytes(np.ascontiguousarray(raw, dtype=np.uint8))
    body = b""
    if junk:
        body += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0"
    fb = _fmt_body(tag, 2, rate, bits, extensible)
    body += b"fmt " + struct.pack("<I", len(fb)) + fb
    dlen = len(raw) if data_len_override is None else data_len_override
    body += b"data" + struct.pack("<I", dlen) + raw + (b"\0" if len(raw) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", (4 + len(body)) & 0xFFFFFFFF) + b"WAVE" + body)


_W64_TAIL = bytes([0xF3, 0xAC, 0xD3, 0x11, 0x8C, 0xD1, 0x00, 0xC0, 0x4F, 0x8E, 0xDB, 0x8A])
_W64_RIFF = b"riff" + bytes([0x2E, 0x91, 0xCF, 0x11, 0xA5, 0xD6, 0x28, 0xDB, 0x04, 0xC1, 0x00, 0x00])


def write_w64(path, raw, fmt, rate):
    """Sony Wave64: 16-byte GUID chunk ids, 64-bit chunk sizes that include the 24-byte header, 8-byte alignment."""
    tag, bits = _tag_bits(fmt)
    raw = bytes(np.ascontiguousarray(raw, dtype=np.uint8))

    def chunk(cid, payload):
        size = 24 + len(payload)
        pad = (-size) % 8
        return cid + _W64_TAIL + struct.pack("<Q", size) + payload + b"\0" * pad

    body = chunk(b"fmt ", _fmt_body(tag, 2, rate, bits)) + chunk(b"data", raw)
    with open(path, "wb") as f:
        f.write(_W64_RIFF + struct.pack("<Q", 40 + len(body)) + b"wave" + _W64_TAIL + body)
