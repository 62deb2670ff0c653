#!/usr/bin/env python
"""Parity of the time-chunk pipeline (FMR_TIME_CHUNKS, FMR_FUSED_CHUNKS) against the default single-chunk schedule
on the same device input: same per-block lengths, audio within the float tolerance, same PLL lock counter."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from airspy_fmradion_b200 import FmDecoder  # noqa: E402


def main():
    import torch
    fs, blk, per, calls, nch = 1.0e7, 2048, 160, 4, 4
    n = blk * per * calls
    # any FM signal does (the comparison is GPU schedule against GPU schedule): mono tone + 19 kHz pilot, some noise
    t = np.arange(n) / fs
    rng = np.random.default_rng(1)
    rows = []
    for c in range(nch):
        mpx = 0.6 * np.sin(2 * np.pi * (1000.0 + 37 * c) * t) + 0.1 * np.sin(2 * np.pi * 19000.0 * t)
        ph = 2 * np.pi * 75000.0 * np.cumsum(mpx) / fs
        x = 0.5 * np.exp(1j * ph) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        rows.append(x.astype(np.complex64))
    iq = np.stack(rows)
    dev = torch.device("cuda", 0)
    d_iq = torch.from_numpy(iq.view(np.float32)).to(dev)
    kw = dict(stereo=True, input_rate=fs, n_channels=nch, max_samples_per_call=blk * per, max_blocks_per_call=per)
    res = {}
    for name, env in (("default", {}), ("tc4_fused", {"FMR_TIME_CHUNKS": "4", "FMR_FUSED_CHUNKS": "1"}),
                      ("tc3_unfused", {"FMR_TIME_CHUNKS": "3"})):
        for k in ("FMR_TIME_CHUNKS", "FMR_FUSED_CHUNKS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        dec = FmDecoder(**kw)
        cap = 8192
        outs, lens = [], []
        st = torch.cuda.current_stream()
        for k in range(calls):
            audio = torch.zeros((nch, cap), dtype=torch.float64, device=dev)
            seg = d_iq[:, 2 * k * blk * per:]
            l = dec.process_device(seg.data_ptr(), n, [blk] * per,
                                   audio.data_ptr(), cap, st.cuda_stream)
            torch.cuda.synchronize()
            outs.append(audio[:, :int(l.sum())].cpu().numpy())
            lens.append(l)
        res[name] = (np.concatenate(outs, axis=1), np.concatenate(lens), dec.stats(0).pll_lock_cnt, dec.last_launches())
        dec.close()
    a0, l0, k0, _ = res["default"]
    ok = True
    for name in ("tc4_fused", "tc3_unfused"):
        a, l, k, nl = res[name]
        d = np.abs(a - a0).max() if a.size else -1.0
        same = list(l) == list(l0) and k == k0 and a.shape == a0.shape
        print("%s: launches %d, lens/lock equal %s, audio %s, max |diff| %.3e" % (name, nl, same, a.shape, d))
        ok &= same and d <= 2e-5
    print("CHUNK PARITY", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
