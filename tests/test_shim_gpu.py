"""The C++ shim (airspy_fmradion_b200/host/fmradion_b200_shim.hpp) used the way main.cpp uses the
reference's FmDecoder: one block per process() call. Compiled with g++ against the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from oracle import siggen
from tests.oracle_select import oracle_fm_run

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_drop_in(tmp_path):
    exe = str(tmp_path / "shim_smoke")
    pkg = os.path.join(ROOT, "airspy_fmradion_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "shim_smoke.cpp"), "-o", exe,
                           "-L" + pkg, "-lfmradion_b200", "-Wl,-rpath," + pkg])
    fs, blk, nblk = 1.0e6, 2048, 150
    iq = siggen.fm_stereo_iq(fs, blk * nblk, 0)
    fin, fout = str(tmp_path / "iq.cf32"), str(tmp_path / "audio.f64")
    iq.tofile(fin)
    out = subprocess.run([exe, "fm", str(fs), fin, fout, str(blk)], capture_output=True, text=True, check=True)
    print(out.stdout.strip())
    audio = np.fromfile(fout, dtype=np.float64)
    ref_audio, _ = oracle_fm_run(iq, fs, blk, stereo=True)
    assert len(audio) == len(ref_audio) > 1000
    assert np.abs(audio - ref_audio).max() <= 2e-5
