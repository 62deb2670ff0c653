"""TEST INFRASTRUCTURE ONLY: deterministic synthetic IQ generators.

FM stereo: L(t)=sin 2pi f_l t, R(t)=sin 2pi f_r t,
  mpx = 0.45(L+R) + 0.45(L-R) sin(2pi 38000 t) + 0.1 sin(2pi 19000 t),
  phi[n] = phi[n-1] + 2pi 75000 mpx[n]/fs,  x[n] = 0.5 exp(j phi[n]) + N(0, sigma^2) per I/Q.
Channel c uses tones (1000+37c, 2500+53c) Hz and noise seed 1234+c.
The noise source is numpy's PCG64 (not std::mt19937): the same bytes are fed to the
reference, the restatement and the GPU path, which is all parity needs.
"""
import numpy as np


def fm_stereo_iq(fs, n, channel=0, sigma=0.01, amp=0.5, start=0, echo=None, mono=False):
    """Return complex64[n] for samples start..start+n-1 of channel `channel`.

    echo: None or (delay_samples, complex_gain) -> x[n] + g*x[n-delay] (static multipath).
    """
    f_l = 1000.0 + 37.0 * channel
    f_r = 2500.0 + 53.0 * channel
    # Generate from t=0 so that the phase integral is exact for any `start`.
    tot = start + n
    t = np.arange(tot, dtype=np.float64) / fs
    left = np.sin(2 * np.pi * f_l * t)
    right = np.sin(2 * np.pi * f_r * t)
    if mono:
        mpx = 0.9 * 0.5 * (left + right)
    else:
        mpx = (0.45 * (left + right)
               + 0.45 * (left - right) * np.sin(2 * np.pi * 38000.0 * t)
               + 0.1 * np.sin(2 * np.pi * 19000.0 * t))
    phi = np.cumsum(2 * np.pi * 75000.0 * mpx / fs)
    x = amp * np.exp(1j * phi)
    if echo is not None:
        d, g = echo
        y = x.copy()
        y[d:] += g * x[:-d]
        x = y
    rng = np.random.Generator(np.random.PCG64(1234 + channel))
    noise = rng.standard_normal((tot, 2), dtype=np.float32) * np.float32(sigma)
    out = np.empty(tot, dtype=np.complex64)
    out.real = x.real.astype(np.float32) + noise[:, 0]
    out.imag = x.imag.astype(np.float32) + noise[:, 1]
    return out[start:]


def am_iq(fs, n, channel=0, sigma=0.005):
    """x = 0.3 (1 + 0.5 sin 2pi 1000 t) + N(0, sigma^2), seed 7+c."""
    t = np.arange(n, dtype=np.float64) / fs
    env = 0.3 * (1.0 + 0.5 * np.sin(2 * np.pi * (1000.0 + 11.0 * channel) * t))
    rng = np.random.Generator(np.random.PCG64(7 + channel))
    noise = rng.standard_normal((n, 2), dtype=np.float32) * np.float32(sigma)
    out = np.empty(n, dtype=np.complex64)
    out.real = env.astype(np.float32) + noise[:, 0]
    out.imag = noise[:, 1]
    return out


def nbfm_iq(fs, n, channel=0, sigma=0.005, amp=0.5, dev=3000.0):
    """Narrow-band FM test signal: a two-tone voice-band message (600 + 37c Hz, 1700 + 53c Hz)
    with +-`dev` Hz peak deviation on a carrier 300 Hz off the centre (so that the tuning-offset
    getter has something to report), + N(0, sigma^2), seed 11+c."""
    t = np.arange(n, dtype=np.float64) / fs
    msg = 0.6 * np.sin(2 * np.pi * (600.0 + 37.0 * channel) * t) + 0.4 * np.sin(2 * np.pi * (1700.0 + 53.0 * channel) * t)
    phi = np.cumsum(2 * np.pi * (300.0 + dev * msg) / fs)
    rng = np.random.Generator(np.random.PCG64(11 + channel))
    noise = rng.standard_normal((n, 2), dtype=np.float32) * np.float32(sigma)
    out = np.empty(n, dtype=np.complex64)
    out.real = (amp * np.cos(phi)).astype(np.float32) + noise[:, 0]
    out.imag = (amp * np.sin(phi)).astype(np.float32) + noise[:, 1]
    return out


def ssb_iq(fs, n, channel=0, sigma=0.003):
    """Test signal for the DSB/USB/LSB/CW/WSPR branches of AmDecoder: complex tones on both sides of
    the carrier (+700 + 17c Hz and +1900 Hz: upper sideband; -1100 Hz: lower sideband; +120 Hz: inside
    the 500 Hz CW filter; +1480 Hz: inside the WSPR band) + N(0, sigma^2), seed 23+c."""
    t = np.arange(n, dtype=np.float64) / fs
    x = (0.30 * np.exp(2j * np.pi * (700.0 + 17.0 * channel) * t) + 0.15 * np.exp(2j * np.pi * 1900.0 * t)
         + 0.20 * np.exp(-2j * np.pi * 1100.0 * t) + 0.25 * np.exp(2j * np.pi * 120.0 * t)
         + 0.10 * np.exp(2j * np.pi * 1480.0 * t))
    rng = np.random.Generator(np.random.PCG64(23 + channel))
    noise = rng.standard_normal((n, 2), dtype=np.float32) * np.float32(sigma)
    out = np.empty(n, dtype=np.complex64)
    out.real = x.real.astype(np.float32) + noise[:, 0]
    out.imag = x.imag.astype(np.float32) + noise[:, 1]
    return out
