"""Host side of SURVEY.md §8 f1 / f4 without a GPU: the C++ FileSource / SndfileOutput of
airspy_fmradion_b200/host/fmradion_b200_io.hpp against the numpy restatement (oracle/fileio.py) and against
Python's own WAV readers (`wave`, scipy.io.wavfile) as an independent reading of the same files."""
import os
import subprocess
import wave

import numpy as np
import pytest

from oracle import fileio, siggen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("io") / "io_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", os.path.join(ROOT, "tests/cpp/io_host_test.cpp"),
                           "-o", out])
    return out


def _run(exe, *args, ok=True):
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=120)
    assert (r.returncode == 0) == ok, r.stdout + r.stderr
    return r.stdout


def _info(line):
    return dict(kv.split("=") for kv in line.split())


def _signal(n, seed=0):
    return siggen.fm_stereo_iq(1.0e6, n, seed)


CASES = [  # (name, writer, fmt, configuration suffix)
    ("wav_s16", "wav", fileio.IQ_S16, ""),
    ("wav_u8", "wav", fileio.IQ_U8, ""),
    ("wav_s24", "wav", fileio.IQ_S24, ""),
    ("wav_f32", "wav", fileio.IQ_CF32, ""),
    ("wavex_s16", "wavex", fileio.IQ_S16, ""),
    ("wavex_f32", "wavex", fileio.IQ_CF32, ""),
    ("wav_junk_s24", "wavjunk", fileio.IQ_S24, ""),
    ("w64_s16", "w64", fileio.IQ_S16, ""),
    ("w64_s24", "w64", fileio.IQ_S24, ""),
    ("raw_s8", "raw", fileio.IQ_S8, ",raw,format=S8_LE,srate=1000k"),
    ("raw_u8", "raw", fileio.IQ_U8, ",raw,format=U8_LE,srate=1000000"),
    ("raw_s16", "raw", fileio.IQ_S16, ",raw,srate=1000000"),  # S16_LE is the default format
    ("raw_s24", "raw", fileio.IQ_S24, ",raw,format=S24_LE,srate=1000000"),
    ("raw_f32", "raw", fileio.IQ_CF32, ",raw,format=FLOAT,srate=1000000"),
]


@pytest.mark.parametrize("name,writer,fmt,suffix", CASES)
def test_filesource_reads_like_sf_read_float(exe, tmp_path, name, writer, fmt, suffix):
    n = 5 * 2048 + 777  # ragged last block
    raw = fileio.quantize_iq(_signal(n, seed=len(name)), fmt)
    path = str(tmp_path / (name + ".bin"))
    if writer == "raw":
        raw.tofile(path)
    elif writer == "w64":
        fileio.write_w64(path, raw, fmt, 1000000)
    else:
        fileio.write_wav(path, raw, fmt, 1000000, extensible=(writer == "wavex"), junk=(writer == "wavjunk"))
    out = str(tmp_path / "out.cf32")
    txt = _run(exe, "read", "filename=%s%s" % (path, suffix), out)
    info = _info(txt.splitlines()[0])
    assert int(info["rate"]) == 1000000 and int(info["fmt"]) == fmt and int(info["total"]) == n
    assert info["container"] == {"raw": "RAW", "w64": "W64", "wavex": "WAVEX"}.get(writer, "WAV")
    assert int(info["low_if"]) == 1 and int(info["blklen"]) == 2048
    assert _info(txt.splitlines()[1]) == {"blocks": "6", "last": "777"}
    got = np.fromfile(out, dtype=np.complex64)
    want = fileio.sf_read_float(raw, fmt)
    assert len(got) == n and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the raw hand-over (device-side decode) delivers the file's own bytes, block by block
    txt = _run(exe, "raw", "filename=%s%s" % (path, suffix), out)
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), raw)


def test_filesource_against_python_wave_readers(exe, tmp_path):
    """Independent reading of the same files: `wave` (PCM) and scipy (float)."""
    from scipy.io import wavfile
    n = 3000
    for fmt, width in ((fileio.IQ_S16, 2), (fileio.IQ_U8, 1), (fileio.IQ_S24, 3)):
        raw = fileio.quantize_iq(_signal(n, seed=fmt), fmt)
        path = str(tmp_path / ("w%d.wav" % width))
        with wave.open(path, "wb") as w:  # written by Python, read by the C++ FileSource
            w.setnchannels(2)
            w.setsampwidth(width)
            w.setframerate(250000)
            w.writeframes(raw.tobytes())
        out = str(tmp_path / "o.cf32")
        info = _info(_run(exe, "read", "filename=" + path, out).splitlines()[0])
        assert int(info["rate"]) == 250000  # "overwrite sample rate" (FileSource.cpp:180-185)
        got = np.fromfile(out, dtype=np.complex64)
        rate, data = wavfile.read(path)  # int16 / uint8 / int32 (24-bit left-justified)
        assert rate == 250000
        if width == 2:
            want = data.astype(np.float64) / 32768.0
        elif width == 1:
            want = (data.astype(np.float64) - 128.0) / 128.0
        else:
            want = data.astype(np.float64) / 2147483648.0
        assert np.array_equal(got.real.astype(np.float64), want[:, 0]) and np.array_equal(got.imag.astype(np.float64), want[:, 1])
    x = _signal(n, 9)
    path = str(tmp_path / "f.wav")
    wavfile.write(path, 384000, np.stack([x.real, x.imag], axis=1).astype(np.float32))
    out = str(tmp_path / "o.cf32")
    _run(exe, "read", "filename=" + path, out)
    assert np.array_equal(np.fromfile(out, dtype=np.complex64), x)


def test_filesource_configuration_and_errors(exe, tmp_path):
    raw = fileio.quantize_iq(_signal(4096), fileio.IQ_S16)
    path = str(tmp_path / "a.wav")
    fileio.write_wav(path, raw, fileio.IQ_S16, 48000)
    out = str(tmp_path / "o.cf32")
    # blklen longer than 10 ms is rounded down to a power of two (FileSource.cpp:236-244): 480 -> 256
    info = _info(_run(exe, "read", "filename=%s,blklen=4096,zero_offset,freq=82500k" % path, out).splitlines()[0])
    assert int(info["blklen"]) == 256 and int(info["low_if"]) == 0 and int(info["freq"]) == 82500000
    info = _info(_run(exe, "read", "filename=%s,blklen=300" % path, out).splitlines()[0])
    assert int(info["blklen"]) == 300
    # streamed WAV with an unknown data length: read to the end of the file
    path2 = str(tmp_path / "b.wav")
    fileio.write_wav(path2, raw, fileio.IQ_S16, 48000, data_len_override=0xFFFFFFFF)
    assert int(_info(_run(exe, "read", "filename=" + path2, out).splitlines()[0])["total"]) == 4096
    assert "Failed to open" in _run(exe, "read", "filename=/nonexistent/x.wav", out, ok=False)
    assert "invalid blklen" in _run(exe, "read", "filename=%s,blklen=0" % path, out, ok=False)
    assert "invalid samplerate" in _run(exe, "read", "filename=%s,srate=abc" % path, out, ok=False)
    assert "not supported" in _run(exe, "read", "filename=%s,format=S32_LE" % path, out, ok=False)
    # PCM_32 is a sub-type FileSource refuses (FileSource.cpp:195-201,342-346)
    import struct
    p32 = str(tmp_path / "c.wav")
    body = b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, 48000, 48000 * 8, 8, 32) + b"data" + struct.pack("<I", 16) + b"\0" * 16
    open(p32, "wb").write(b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WAVE" + body)
    assert "Unsupported sub type" in _run(exe, "read", "filename=" + p32, out, ok=False)
    pbad = str(tmp_path / "d.flac")
    open(pbad, "wb").write(b"fLaC" + b"\0" * 64)
    assert "Unsupported major format" in _run(exe, "read", "filename=" + pbad, out, ok=False)


def test_helpers(exe):
    txt = _run(exe, "misc").splitlines()
    assert txt[0] == "map: [alpha]=[100] [beta]=[] [delta]=[] [gamma]=[x=yz]"
    assert txt[1] == "int: 1 10000000 0 0 0 1 -42"
    assert txt[2] == "round_power: 0 1 256 4096"
    assert txt[3] == "pps: [       3        1234567  1700000000.250000   -12.346]"  # "{:>8} {:>14} {:18.6f} {:+9.3f}"
    assert txt[4] == "squelch: 0.01 0"
    f = txt[5].split()
    assert f[1:4] == ["0", "1", "1"]
    if_level = np.float32(0.75 * float(np.float32(0.25 * 0.5)) + 0.25 * 0.5)
    assert abs(float(f[4]) - float(if_level)) < 1e-7 and abs(float(f[5]) - 0.05 * 0.25) < 1e-8


@pytest.mark.parametrize("kind", ["wav16", "wavf32", "raw16", "rawf32"])
def test_sndfile_output_writes_like_sf_write_double(exe, tmp_path, kind):
    from scipy.io import wavfile
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1.0, 1.0, 10000), [0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.5 / 32767, 1.0, -1.0]])
    x = x[: len(x) // 2 * 2]
    fin, fout = str(tmp_path / "x.f64"), str(tmp_path / ("y." + kind))
    x.tofile(fin)
    txt = _run(exe, "write", fin, fout, "48000", "1", kind, "1000")
    assert "written=%d" % len(x) in txt
    out_fmt = fileio.OUT_S16 if kind.endswith("16") else fileio.OUT_F32
    want = fileio.sf_write_double(x, out_fmt)
    if kind.startswith("raw"):
        got = np.fromfile(fout, dtype=want.dtype)
    else:
        rate, data = wavfile.read(fout)  # header patched after every write (SFC_SET_UPDATE_HEADER_AUTO)
        assert rate == 48000 and data.shape == (len(x) // 2, 2) and data.dtype == want.dtype
        got = data.reshape(-1)
    assert np.array_equal(got, want)
    if out_fmt == fileio.OUT_S16:  # full scale: +1.0 -> 32767, -1.0 -> -32767 (no clipping needed, no wrap)
        assert list(want[-2:]) == [32767, -32767]


def test_output_stage_restatement():
    a = [np.array([0.5, -0.5, 0.25, 0.25]), np.zeros(0), np.array([1.0, -1.0])]
    out, lv, if_level, audio_level = fileio.output_stage(a, [0.2, None, 0.001], fileio.OUT_S16, squelch_level=0.01)
    assert list(out) == [8192, -8192, 4096, 4096, 0, 0]  # rint(0.25 * 32767) = 8192 (8191.75), second block muted
    assert lv[1, 0] == -1 and lv[2, 3] == 0 and lv[0, 3] == 0.5
    assert abs(lv[0, 2] - np.sqrt((0.25 + 0.25 + 0.0625 + 0.0625) / 4)) < 1e-7
