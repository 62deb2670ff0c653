"""Host-compiled checks of device-side bookkeeping that can be verified without a GPU: the sources
under csrc/ are written so that the arithmetic core also compiles with g++ (FMR_HD)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(src, tmp_path, extra=()):
    exe = str(tmp_path / os.path.splitext(os.path.basename(src))[0])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", *extra, "-o", exe, os.path.join(ROOT, src)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_streaming_halfband_bookkeeping(tmp_path):
    """HbsCascade (register delay lines, warm-up length, stage delays) == direct three-stage
    evaluation, bit for bit, from NaN-poisoned state, for every tile alignment and block factor."""
    out = _build_and_run("tests/cpp/hbstream_host_test.cpp", tmp_path)
    assert "FAIL" not in out and out.count(": ok") >= 5


def test_branch_free_fast_atan2(tmp_path):
    out = _build_and_run("tests/cpp/fast_atan2_host_test.cpp", tmp_path, extra=("-frounding-math",))
    assert "exact-division form mismatches=0" in out


def test_inplace_fft_host_emulation(tmp_path):
    """k_fir_fft_ip's per-thread pass bodies (fmr_fft_inplace.cuh, __host__ __device__) run on the CPU, threads of a
    pass in scrambled order, against the direct circular convolution in double; pad words must stay untouched."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    exe = str(tmp_path / "fft_inplace_host_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-arch=sm_100a", "-x", "cu",
                           os.path.join(ROOT, "tests/cpp/fft_inplace_host_test.cu"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and "inplace fft: ok" in out.stdout
