"""Frequency-domain resampler (csrc/fmr_fdr.cuh), CPU side.

1. The two properties of r8brain's design that make the 2307-tap low-pass + 192 x 18 polyphase bank
   (CDSPBlockConvolver.h:252-353, CDSPFracInterpolator.h:861-925) equal to band-limited resampling, checked on the
   frozen tables of every 625:192 chain: the low-pass is symmetric and below 2e-9 from the output Nyquist on, and every
   bank row is an ideal fractional delay of 8 + p/192 samples to 1e-8 over the band the low-pass leaves.
2. The kernel's per-thread pass bodies, compiled for the host, against the reference's two stages in double."""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "airspy_fmradion_b200", "csrc", "fmr_tables_generated.inc")


def _arr(text, name):
    m = re.search(r"static const double %s\[(\d+)\] = \{(.*?)\};" % name, text, re.S)
    return np.array([float(x) for x in m.group(2).replace("\n", " ").split(",") if x.strip()])


def test_tables_allow_band_limited_resampling():
    text = open(INC).read()
    # chain index in fmr_tables_generated.inc -> (instep, outstep, flen): every chain the frequency-domain form serves
    chains = {"chain0": (625, 192, 18), "chain2": (625, 192, 18),   # 10 MHz, 2.5 MHz -> 384 kHz
              "chain1": (125, 32, 18), "chain6": (125, 32, 18),     # 6 MHz, 3 MHz -> 384 kHz
              "chain3": (125, 48, 24), "chain18": (125, 48, 24)}    # 1 MHz -> 384 kHz, 1 MHz -> 48 kHz
    for chain, (instep, outstep, flen) in chains.items():
        bc, fi = _arr(text, chain + "_bc"), _arr(text, chain + "_fi").reshape(outstep, flen)
        assert np.array_equal(bc, bc[::-1]), chain
        assert (len(bc) - 1) // 2 + flen // 2 + 1 <= 1500  # guard of a 10000-sample block
        n = 1 << 18
        mag = np.abs(np.fft.rfft(np.concatenate([bc, np.zeros(n - len(bc))]))) / bc.sum()
        nyq = 0.5 * outstep / instep  # output Nyquist on the input axis = the edge of the kept band
        edge = int(np.ceil(nyq * n))
        assert mag[edge:].max() < 2e-9, (chain, mag[edge:].max())
        f = np.linspace(0.0, nyq * 1.04, 321)
        e = np.exp(-2j * np.pi * np.outer(f, np.arange(flen)))
        worst = 0.0
        for p in range(outstep):
            # row p evaluates the stream (flen / 2 - 1) + frac(p * instep / outstep) samples into its window
            delay = (flen // 2 - 1) + ((p * instep) % outstep) / outstep if chain in ("", ) else None
            h = e @ fi[p]
            ph = np.unwrap(np.angle(h))
            d = -(ph[8] - ph[0]) / (2 * np.pi * (f[8] - f[0]))
            assert abs(d - (flen // 2 - 1) - p / outstep) < 1e-6, (chain, p, d)  # bank row p <-> fraction p / outstep
            worst = max(worst, np.abs(h * np.exp(2j * np.pi * f * ((flen // 2 - 1) + p / outstep)) - 1.0).max())
        assert worst < 3e-8, (chain, worst)


def test_fdr_host_emulation(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    exe = str(tmp_path / "fdr_host_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-arch=sm_100a", "-x", "cu", "-w",
                           os.path.join(ROOT, "tests/cpp/fdr_host_test.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and "fdr host emulation: ok" in out.stdout
