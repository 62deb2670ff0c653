#!/usr/bin/env python
"""Device-resident A/B sweep of kernel variants in ONE process (the synthetic IQ is generated
once): every variant = environment switches read at handle creation + a step size.
Prints one JSON line per variant; not a bench line (bench.py is the contract)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

VARIANTS = [
    # name, env, blocks per step, channels
    ("all_on_329", {}, 329, 8192),
    ("all_on_128", {}, 128, 8192),
    ("all_on_128_c1024", {}, 128, 1024),
    ("all_on_128_c148", {}, 128, 148),
    # time-chunk pipeline with the fused 384 kHz core (front end of chunk k+1 overlaps core + audio tail of chunk k)
    ("tc2_fused", {"FMR_TIME_CHUNKS": "2", "FMR_FUSED_CHUNKS": "1"}, 329, 8192),
    ("tc3_fused", {"FMR_TIME_CHUNKS": "3", "FMR_FUSED_CHUNKS": "1"}, 329, 8192),
    ("tc4_fused", {"FMR_TIME_CHUNKS": "4", "FMR_FUSED_CHUNKS": "1"}, 329, 8192),
    ("tc6_fused", {"FMR_TIME_CHUNKS": "6", "FMR_FUSED_CHUNKS": "1"}, 329, 8192),
    ("tc4_unfused", {"FMR_TIME_CHUNKS": "4"}, 329, 8192),
    ("tc4_fused_c2048", {"FMR_TIME_CHUNKS": "4", "FMR_FUSED_CHUNKS": "1"}, 329, 2048),
    ("all_on_329_c2048", {}, 329, 2048),
    ("fft_tw_off", {"FMR_FFT_TW": "0"}, 329, 8192),
    ("fft_1024thr", {"FMR_FFT_THREADS": "1024"}, 329, 8192),
    ("fft_inplace_512", {"FMR_FFT_INPLACE": "1"}, 329, 8192),
    ("fft_inplace_1024", {"FMR_FFT_INPLACE": "1", "FMR_FFT_THREADS": "1024"}, 329, 8192),
    ("fft_stockham", {"FMR_FFT_INPLACE": "0"}, 329, 8192),
    ("fft_inplace_r32", {"FMR_FFT_INPLACE": "2"}, 329, 8192),
    ("fft_epi_smem", {"FMR_FFT_EPI": "1"}, 329, 8192),
    ("fft_inplace_r32_epi", {"FMR_FFT_INPLACE": "2", "FMR_FFT_EPI": "1"}, 329, 8192),
    ("fft_inplace_8k", {"FMR_FFT_INPLACE8K": "1"}, 329, 8192),
    ("fdr_off", {"FMR_FDR": "0"}, 329, 8192),
    ("fe_off", {"FMR_FE": "0"}, 329, 8192),
    ("fe_a", {"FMR_FE_VARIANT": "0"}, 329, 8192),
    ("fe_b", {"FMR_FE_VARIANT": "1"}, 329, 8192),
    ("fe_c", {"FMR_FE_VARIANT": "2"}, 329, 8192),
    ("fe_s", {"FMR_FE_VARIANT": "3"}, 329, 8192),
    ("fe_a_658", {"FMR_FE_VARIANT": "0"}, 658, 8192),
    ("all_on_329_c16384", {}, 329, 16384),
    ("fft_inplace_r32_8k", {"FMR_FFT_INPLACE": "2", "FMR_FFT_INPLACE8K": "1"}, 329, 8192),
]


def main():
    import torch
    only = set(sys.argv[1:])
    wl = "cfg2_fm_stereo_10Msps"
    fs, stereo, mpf, mode = bench.WORKLOADS[wl]
    variants = [v for v in VARIANTS if not only or v[0] in only]
    Cmax = max(v[3] for v in variants)
    Tmax = max(v[2] for v in variants) * bench.BLK
    if Cmax * Tmax * 8 > 60e9:  # keep the input under 60 GB: fewer channels for the long steps
        Tmax_c = int(60e9 / 8 / Cmax) // bench.BLK * bench.BLK
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    # one [Cgen, Tmax] array; variants with more channels reuse rows (row stride stays Tmax)
    Cgen = min(Cmax, int(60e9 / 8 / Tmax))
    iq = bench.gen_iq_device(torch, dev, fs, Cgen, Tmax, mode)
    stream = torch.cuda.current_stream()
    for name, env, nblk, C in variants:
        for k in ("FMR_HB_STREAM", "FMR_FUSE_FI", "FMR_FFT_F64", "FMR_HBS_TILE", "FMR_FFT", "FMR_SERIAL_V2", "FMR_TIME_CHUNKS", "FMR_SERIAL_SMS", "FMR_CORE_FUSED", "FMR_HBS_STAGES", "FMR_HBS_L2PF", "FMR_HBS_TMA", "FMR_FFT_N", "FMR_CORE_ROT", "FMR_FUSED_CHUNKS", "FMR_CHUNK_MIN_BLOCKS", "FMR_FFT_TW", "FMR_FFT_THREADS", "FMR_FFT_INPLACE", "FMR_FFT_INPLACE8K", "FMR_FFT_REGCAP", "FMR_FFT_EPI", "FMR_FDR", "FMR_FE", "FMR_FE_VARIANT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        C = min(C, Cgen)
        T = nblk * bench.BLK
        dec = bench.make_decoder(wl, C, T, nblk, 0)
        audio_cap = int(T * 48000.0 / fs) * 2 + 64
        audio = torch.zeros((C, audio_cap), dtype=torch.float64, device=dev)
        bl = [bench.BLK] * nblk

        def step():
            return dec.process_device(iq.data_ptr(), Tmax, bl, audio.data_ptr(), audio_cap, stream.cuda_stream)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        dec.set_profiling(True)
        stage = {}
        for _ in range(2):
            step()
            torch.cuda.synchronize()
            for k, v in dec.stage_times().items():
                stage[k] = stage.get(k, 0.0) + v / 2
        dec.set_profiling(False)
        print(json.dumps({"variant": name, "env": env, "channels": C, "blocks": nblk, "ms_per_step": round(ms, 3),
                          "gsamples_per_s": round(C * T / ms / 1e6, 2), "launches": dec.last_launches(),
                          "checksum": float(audio[:, :1000].abs().sum().item()),
                          "stage_ms": {k: round(v, 3) for k, v in stage.items()}}), flush=True)
        dec.close()
        del audio


if __name__ == "__main__":
    main()
