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

# every environment switch the library still reads (all select between paths the GPU suite covers)
SWITCHES = ("FMR_FE", "FMR_FE_VARIANT", "FMR_FE_MIN_BLOCKS", "FMR_FDR", "FMR_FFT_INPLACE", "FMR_FFT_TW", "FMR_FFT_F64", "FMR_FFT",
            "FMR_FUSE_FI", "FMR_HB_STREAM", "FMR_HBS_TMA", "FMR_CORE_FUSED", "FMR_AM_FFT_FILTER", "FMR_AUDIO_FP64")

VARIANTS = [
    # name, env, blocks per step, channels
    ("all_on_329", {}, 329, 8192),
    ("all_on_128", {}, 128, 8192),
    ("all_on_329_c16384", {}, 329, 16384),
    ("all_on_329_c2048", {}, 329, 2048),
    ("all_on_329_c148", {}, 329, 148),
    ("fe_off", {"FMR_FE": "0"}, 329, 8192),             # unfused front end: k_hb_stream_tma -> ring -> k_fdr
    ("fe_split", {"FMR_FE_VARIANT": "1"}, 329, 8192),   # fused front end, real / imaginary part on two lanes
    ("fe_cons8", {"FMR_FE_VARIANT": "2"}, 329, 8192),   # fused front end, eight consumer warps + setmaxnreg (CfgA)
    ("fdr_off", {"FMR_FE": "0", "FMR_FDR": "0"}, 329, 8192),  # time-domain form: 16384-point FFT low-pass + bank
    ("fft_stockham", {"FMR_FE": "0", "FMR_FDR": "0", "FMR_FFT_INPLACE": "0"}, 329, 8192),
    ("hb_tiled", {"FMR_FE": "0", "FMR_HB_STREAM": "0"}, 329, 8192),
    ("hbs_cp_async", {"FMR_FE": "0", "FMR_HBS_TMA": "0"}, 329, 8192),
    ("core_unfused", {"FMR_CORE_FUSED": "0"}, 329, 8192),
    ("audio_fp64", {"FMR_AUDIO_FP64": "1"}, 329, 8192),  # audio half-bands, low-pass and pilot cut in FP64
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
        for k in SWITCHES:
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
