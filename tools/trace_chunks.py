#!/usr/bin/env python
"""Timeline of the time-chunk pipeline (FMR_TRACE=1): which launch groups of which chunk overlap.
usage: trace_chunks.py <channels> <time chunks> [extra ENV=VALUE ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    C, tc = int(sys.argv[1]), int(sys.argv[2])
    os.environ.update({"FMR_TRACE": "1", "FMR_TIME_CHUNKS": str(tc), "FMR_FUSED_CHUNKS": "1"})
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        os.environ[k] = v
    wl = "cfg2_fm_stereo_10Msps"
    fs, stereo, mpf, mode = bench.WORKLOADS[wl]
    nblk = 329
    T = nblk * bench.BLK
    dev = torch.device("cuda", 0)
    iq = bench.gen_iq_device(torch, dev, fs, C, T, mode)
    dec = bench.make_decoder(wl, C, T, nblk, 0)
    cap = int(T * 48000.0 / fs) * 2 + 64
    audio = torch.zeros((C, cap), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream()
    for i in range(3):
        sys.stderr.write("---- call %d\n" % i)
        sys.stderr.flush()
        dec.process_device(iq.data_ptr(), T, [bench.BLK] * nblk, audio.data_ptr(), cap, st.cuda_stream)
        torch.cuda.synchronize()
    dec.close()


if __name__ == "__main__":
    main()
