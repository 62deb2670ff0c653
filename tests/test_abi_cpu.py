"""The C-ABI library loads and exports every symbol include/fmradion_b200.h declares; without a
GPU every computing entry point fails loudly (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "fmradion_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(fmr_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported():
    from airspy_fmradion_b200 import _capi
    L = _capi.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    assert sorted(_capi.EXPORTS) == syms, "python binding list and header disagree"
    assert b"sm_100a" in L.fmr_version()


def test_library_is_sm100a_only():
    """The shipped library carries sm_100a SASS (not a generic PTX fallback build)."""
    from airspy_fmradion_b200 import _capi
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from airspy_fmradion_b200 import FmDecoder, FmrError
    with pytest.raises(FmrError) as e:
        FmDecoder(input_rate=1e6)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    """Nothing under airspy_fmradion_b200/ may reference oracle/ (the checker is not the product)."""
    pkg = os.path.join(ROOT, "airspy_fmradion_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "oracle/_ref/libfmref.so", ""), fn


def test_host_schedule_matches_reference_counts():
    """fmr_fm_schedule / fmr_am_schedule (pure host code) reproduce the reference's per-call sizes."""
    from airspy_fmradion_b200 import _capi
    from tests import golden_util as gu
    L = _capi.lib()
    g = gu.golden()
    for name, (mode, fs, nblk, blk, skw, dkw) in gu.CASES.items():
        bl = np.full(nblk, blk, dtype=np.uint32)
        out = np.zeros(nblk, dtype=np.uint32)
        if mode == "fm":
            ifl = np.zeros(nblk, dtype=np.uint32)
            _capi.check(L.fmr_fm_schedule(fs, int(dkw.get("stereo", True)), 0, bl.ctypes.data, nblk, ifl.ctypes.data,
                                          out.ctypes.data))
        else:
            _capi.check(L.fmr_am_schedule(fs, 0, bl.ctypes.data, nblk, out.ctypes.data))
        assert list(out) == list(g[name + "/lens"]), name
    # resumed from the middle of a stream: same counts as the tail of the from-zero schedule
    fs, nblk, blk = 1.0e7, 500, 2048
    bl = np.full(nblk, blk, dtype=np.uint32)
    full = np.zeros(nblk, dtype=np.uint32)
    _capi.check(L.fmr_fm_schedule(fs, 1, 0, bl.ctypes.data, nblk, None, full.ctypes.data))
    tail = np.zeros(nblk - 123, dtype=np.uint32)
    _capi.check(L.fmr_fm_schedule(fs, 1, 123 * blk, bl.ctypes.data, nblk - 123, None, tail.ctypes.data))
    assert list(tail) == list(full[123:])
    # unsupported rate is reported, not guessed
    assert L.fmr_fm_schedule(1234567.0, 1, 0, bl.ctypes.data, 1, None, full.ctypes.data) == 2
