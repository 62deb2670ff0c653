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
    for chain in ("chain0", "chain2"):  # 10 MHz and 2.5 MHz -> 384 kHz: 1.25 MHz in front of the low-pass
        bc, fi = _arr(text, chain + "_bc"), _arr(text, chain + "_fi").reshape(192, 18)
        assert len(bc) == 2307 and np.array_equal(bc, bc[::-1])
        n = 1 << 18
        mag = np.abs(np.fft.rfft(np.concatenate([bc, np.zeros(n - len(bc))]))) / bc.sum()
        edge = int(np.ceil(192.0 / 1250.0 * n))  # output Nyquist (192 kHz) on the 1.25 MHz axis
        assert mag[edge:].max() < 2e-9
        f = np.linspace(0.0, 0.16, 321)
        e = np.exp(-2j * np.pi * np.outer(f, np.arange(18)))
        worst = 0.0
        for p in range(192):
            h = e @ fi[p]
            worst = max(worst, np.abs(h * np.exp(2j * np.pi * f * (8 + p / 192.0)) - 1.0).max())
        assert worst < 2e-8, worst


def test_fdr_host_emulation(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    exe = str(tmp_path / "fdr_host_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-arch=sm_100a", "-x", "cu", "-w",
                           os.path.join(ROOT, "tests/cpp/fdr_host_test.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and "fdr host emulation: ok" in out.stdout
