#!/usr/bin/env python
"""Ten-second sanity check of the built library without torch: one 10 Msps FM handle, two calls, finite audio."""
import sys, os, time
t0 = time.time()
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from airspy_fmradion_b200 import FmDecoder
fs, blk, per = 1.0e7, 2048, 96
rng = np.random.default_rng(0)
n = blk * per
ph = np.cumsum(0.05 * np.sin(2 * np.pi * 1000.0 * np.arange(2 * n) / fs))
iq = (0.5 * np.exp(1j * ph)).astype(np.complex64)
dec = FmDecoder(stereo=True, input_rate=fs, n_channels=2, max_samples_per_call=n, max_blocks_per_call=per)
tot = 0
for k in range(2):
    a, l = dec.process_blocks(np.stack([iq[k * n:(k + 1) * n]] * 2), [blk] * per)
    tot += a.shape[1]
    assert np.isfinite(a).all()
print("quick check ok: %d audio doubles, launches %d, %.1f s" % (tot, dec.last_launches(), time.time() - t0))
