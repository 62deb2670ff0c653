#!/usr/bin/env python
"""bench.py — throughput of the demodulation hot path on B200 (and of the reference on CPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path
    torchrun ... bench.py --gpus N ...                       # one rank per GPU

A "step" is one pass of the hot path over one batch of synthetic IQ: every channel of the
handle consumes `--blocks` source blocks of 2048 samples (FileSource's default block,
FileSource.h:34) handed over as one super-block with its block partition. The default of 329
blocks (673 792 samples, 67 ms of signal per channel) fills one 8192-point block of the audio
resampler's low-pass per channel and step; the IF resampler's 10000-point blocks lie on an absolute
grid, ~11 per step, of which the fused front end takes those that lie inside the step.
Streams are continuous across steps (filter/PLL/AGC state carries over).
Metric = IQ Msamples/s consumed, summed over channels and GPUs (BASELINE.json).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (input_rate, stereo, multipath_stages, mode)
    "cfg2_fm_stereo_10Msps": (1.0e7, True, 0, "fm"),
    "cfg3_fm_stereo_10Msps_E200": (1.0e7, True, 200, "fm"),
    "cfg4_fm_stereo_1Msps": (1.0e6, True, 0, "fm"),
    "cfg5_am_384ksps": (384000.0, False, 0, "am"),
}
BLK = 2048


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t_start=None, t_end=None):
        """Clocks over [t_start, t_end] (the timed region; the sampler itself is started before the warm-up so that
        nvidia-smi is already polling when it begins). With no sample inside the window the nearest ones are used and
        `samples_in_window` says 0."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if t_start is None or (t_start - 0.05 <= r[0] <= t_end + 0.05)]
        n_in = len(inside)
        if not inside and rows and t_end is not None:
            inside = sorted(rows, key=lambda r: abs(r[0] - t_end))[:2]
        sm, mx, reasons = [], [], set()
        for _, a, b, flags in inside:
            sm.append(a)
            mx.append(b)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_window": n_in}


def gen_iq_device(torch, dev, fs, C, T, mode, chunk=32):
    """Synthetic IQ on the device, [C, T] complex64 (SURVEY.md §8(d) signal model; torch RNG)."""
    out = torch.empty((C, T), dtype=torch.complex64, device=dev)
    t = torch.arange(T, dtype=torch.float64, device=dev) / fs
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    for c0 in range(0, C, chunk):
        c1 = min(C, c0 + chunk)
        ch = torch.arange(c0, c1, dtype=torch.float64, device=dev)[:, None]
        if mode == "fm":
            L = torch.sin(2 * np.pi * (1000.0 + 37.0 * ch) * t)
            R = torch.sin(2 * np.pi * (2500.0 + 53.0 * ch) * t)
            mpx = 0.45 * (L + R) + 0.45 * (L - R) * torch.sin(2 * np.pi * 38000.0 * t) \
                + 0.1 * torch.sin(2 * np.pi * 19000.0 * t)
            del L, R
            phi = torch.cumsum(2 * np.pi * 75000.0 * mpx / fs, dim=1)
            del mpx
            re, im, sigma = 0.5 * torch.cos(phi), 0.5 * torch.sin(phi), 0.01
            del phi
        else:
            re = 0.3 * (1.0 + 0.5 * torch.sin(2 * np.pi * (1000.0 + 11.0 * ch) * t))
            im, sigma = torch.zeros_like(re), 0.005
        noise = torch.randn((c1 - c0, T, 2), dtype=torch.float32, device=dev, generator=g) * sigma
        out[c0:c1] = torch.complex(re.float() + noise[..., 0], im.float() + noise[..., 1])
        del re, im, noise
    return out


def make_decoder(wl, C, T, nblk, device):
    from airspy_fmradion_b200 import AmDecoder, FmDecoder
    fs, stereo, mpf, mode = WORKLOADS[wl]
    if mode == "fm":
        return FmDecoder(stereo=stereo, multipath_stages=mpf, input_rate=fs, n_channels=C,
                         max_samples_per_call=T, max_blocks_per_call=nblk, device=device)
    return AmDecoder(input_rate=fs, n_channels=C, max_samples_per_call=T, max_blocks_per_call=nblk, device=device)


def cpu_reference(wl, seconds, threads):
    """Reference CPU path (oracle/_ref = the reference's own classes) on the host cores."""
    from oracle import ref, siggen
    fs, stereo, mpf, mode = WORKLOADS[wl]
    if not ref.available():
        return None
    n_iq = BLK * 2048
    iq = siggen.fm_stereo_iq(fs, n_iq, 0) if mode == "fm" else siggen.am_iq(fs, n_iq, 0)
    m = 0 if mode == "fm" else 1
    # calibrate (1 thread and all threads), then size both samples for ~`seconds` of wall time
    probe = 400 if mpf == 0 else 100
    t = ref.bench(m, fs, stereo, mpf, 1, probe, BLK, iq)
    blocks1 = max(200, int(probe / t * seconds / 2))
    t1 = ref.bench(m, fs, stereo, mpf, 1, blocks1, BLK, iq)
    one = blocks1 * BLK / t1 / 1e6
    t = ref.bench(m, fs, stereo, mpf, threads, probe, BLK, iq)
    blocks = max(200, int(probe / t * seconds))
    tn = ref.bench(m, fs, stereo, mpf, threads, blocks, BLK, iq)
    allc = threads * blocks * BLK / tn / 1e6
    return {"value": allc, "unit": "Msamples/s", "cores": threads, "kind": "reference",
            "one_core_value": one,
            "sample": "%d threads x %d blocks of %d IQ samples (%s), reference classes compiled from "
                      "/root/reference with a generic-C VOLK shim, unthrottled" % (threads, blocks, BLK, wl),
            "seconds": tn}

def rooflines(wl, dec, stage, C, T, n_audio, width, step_ms, peak, peak_src, fs):
    """The `roofline` object of the bench line. Headline = the kernel that streams the IQ input from HBM, on ITS OWN
    algorithmic bytes and its own device time (CUDA events recorded by the library around its launch, on the launching
    stream); `whole_step` = the step's algorithmic bytes (IQ in + audio out) over the step time; `stages` lists every
    stage with the bytes it moves algorithmically, what bounds it, and (HBM-bound stages only) a fraction of the peak."""
    if not stage:
        return None
    fs_mode = WORKLOADS[wl][3]
    plan = dec.last_plan() if hasattr(dec, "last_plan") else {"fused_blocks": 0, "unfused_blocks": 0,
                                                              "unfused_halfband_outputs": 0, "block_in": 0, "block_out": 0}
    n_if = int(T * (384000.0 if fs_mode == "fm" else 48000.0) / fs)  # IF-rate samples per channel per step
    n48 = n_audio // width
    dec_hb = plan["block_in"] // 7500 if plan["block_in"] else 8       # input samples per half-band output
    own = {  # stage -> (bytes per step, bound)
        "if_frontend_fused": (C * plan["fused_blocks"] * (plan["block_in"] + plan["block_out"]) * 8, "hbm"),
        "if_halfband_cascade": (C * (plan["unfused_halfband_outputs"] * (dec_hb + 1) * 8 if plan["block_in"] else T * 9), "hbm"),
        "if_lowpass": (C * (plan["unfused_blocks"] * (10000 + plan["block_out"]) * 8 if plan["block_in"] else (T // 8 + n_if) * 8),
                       "shared memory / FP32 issue (FFT)"),
        "fm_core_fused": (C * n_if * (8 + 16), "latency (serial AGC / PLL recurrences, one lane per channel)"),
        "fm_multipath": (C * n_if * 16, "FP32 issue (sample-serial NLMS)"),
        "audio_halfband_cascade": (C * (n_if * 16 + n_if // 4 * 16), "shared memory / FP64"),
        "audio_lowpass": (C * (n_if // 4 + n48) * 16, "shared memory / FP64 (FFT)"),
        "pilot_cut_fir": (C * n48 * 32, "FP64"),
        "dcblock_matrix": (C * n48 * (16 + 8 * width), "latency (serial IIR, one lane per channel)"),
    }
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
    except Exception:
        pass
    stages = {}
    for k, v in stage.items():
        b, bound = own.get(k, (None, "-"))
        e = {"ms": round(v, 4), "bound": bound}
        if b:
            e["algorithmic_bytes"] = int(b)
            e["gb_per_s"] = round(b / (v * 1e-3) / 1e9, 1)
            if bound == "hbm":
                e["frac"] = round(e["gb_per_s"] / peak, 4)
        stages[k] = e
    head = "if_frontend_fused" if stage.get("if_frontend_fused") else "if_halfband_cascade"
    hb, _ = own[head]
    ach = hb / (stage[head] * 1e-3) / 1e9
    tr = traffic.get(head, {})
    alg_step = C * T * 8 + C * n_audio * 8
    whole = alg_step / (step_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": head, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": (hb * tr["dram_bytes_per_algorithmic_byte"]) if "dram_bytes_per_algorithmic_byte" in tr else None,
            "traffic_source": tr.get("source"),
            "algorithmic_bytes": int(hb), "kernel_ms": stage[head], "peak_source": peak_src,
            "whole_step": {"algorithmic_bytes": int(alg_step), "achieved": whole, "frac": whole / peak, "ms": step_ms},
            "whole_step_frac": whole / peak,
            "front_end_plan": plan, "stages": stages}
    mpf = WORKLOADS[wl][2]
    if mpf > 0 and stage.get("fm_multipath"):
        # the multipath configuration is compute bound, report it against the FP32 roofline too.
        # MultipathFilter.cpp:108-161: 4 * stages + 1 complex taps; per 384 kHz sample one complex FIR (8 flop per tap),
        # every 4th in-call sample the NLMS update (another ~10 flop per tap)
        ntaps = 4 * mpf + 1
        flop = C * n_if * (8 * ntaps + 10 * ntaps / 4.0)
        pk = 148 * 128 * 2 * 1.965e9 / 1e12
        a = flop / (stage["fm_multipath"] * 1e-3) / 1e12
        roof["fp32"] = {"bound": "fp32", "kernel": "fm_multipath (k_mpf)", "achieved": a, "peak": pk, "unit": "TFLOP/s",
                        "frac": a / pk, "flop_per_if_sample": 8 * ntaps + 10 * ntaps / 4.0,
                        "peak_source": "nominal: 148 SMs x 128 FP32 lanes x 2 x 1.965 GHz (no measured FP32 figure in MEASURED_PEAKS.json)"}
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2_fm_stereo_10Msps", choices=sorted(WORKLOADS))
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU (default per workload)")
    ap.add_argument("--blocks", type=int, default=329, help="source blocks of 2048 samples per channel per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--handles", type=int, default=1, metavar="G",
                    help="opt-in: also time the same channels split over G independent handles per GPU, each on its own "
                         "stream (reported as multi_handle; the headline value stays the single-handle figure)")
    ap.add_argument("--single-homed", type=int, default=0, metavar="CH_PER_GPU",
                    help="opt-in, N>1: also time the single-homed I/O mode (int16 IQ of all channels enters at rank 0, "
                         "is scattered over NCCL/NVLink, decoded on the owning GPU, int16 audio gathered back)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    fs, stereo, mpf, mode = WORKLOADS[wl]
    metric = "IQ Msamples/s (%s)" % wl

    nblk = args.blocks
    T = nblk * BLK
    # channels per GPU: many (the serial recurrences cost the same for 1 or 14208 channels, so throughput grows with the
    # channel count, and the input must exceed the L2), fitting 180 GB with their rings, and a multiple of
    # 148 SMs x 32 channels per CTA of the lane-per-channel kernels, so that every SM holds the same number of their
    # CTAs (3 for cfg2; with 16384 channels some SMs hold 4 and the 384 kHz core waits for those: 302 instead of
    # 310 Gsamples/s). cfg4's 384 kHz rings are 10x larger per input sample; cfg3: 592 CTAs of 3 channels = one wave of
    # the multipath kernel (4 CTAs/SM)
    C = args.channels or {"cfg2_fm_stereo_10Msps": 14208, "cfg3_fm_stereo_10Msps_E200": 1776,
                          "cfg4_fm_stereo_1Msps": 9472, "cfg5_am_384ksps": 9472}[wl]
    Ce = min(C, 1024)  # channels of the end-to-end (host buffer) measurement: all of them would need 77 GB of pinned memory
    config = {"workload": wl, "channels_per_gpu": C, "samples_per_channel_per_step": T,
              "block": BLK, "blocks_per_step": nblk, "input_bytes_per_step": C * T * 8,
              "l2": "input per step (%.0f MB) exceeds the 126 MB L2" % (C * T * 8 / 1e6),
              "parallelism": "channels sharded %d per GPU, no data-path collective" % C,
              "e2e_channels_per_gpu": Ce}


    if args.impl == "reference":
        # The reference's own CPU implementation of the path (oracle/_ref) on all host threads. One step = one bounded
        # sample of the same workload (every thread decodes its own stream); W warm-up samples, then K timed ones.
        if rank != 0:
            return 0
        from oracle import ref, siggen
        if not ref.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libfmref.so was not built"}))
            return 0
        threads = os.cpu_count() or 1
        m = 0 if mode == "fm" else 1
        n_iq = BLK * 2048
        iq = siggen.fm_stereo_iq(fs, n_iq, 0) if mode == "fm" else siggen.am_iq(fs, n_iq, 0)
        probe = 400 if mpf == 0 else 100
        t = ref.bench(m, fs, stereo, mpf, threads, probe, BLK, iq)
        blocks = max(200, int(probe / t * max(0.5, args.cpu_seconds / 4)))  # ~1.5 s of wall time per step
        for _ in range(max(0, args.warmup - 1)):
            ref.bench(m, fs, stereo, mpf, threads, blocks, BLK, iq)
        secs = [ref.bench(m, fs, stereo, mpf, threads, blocks, BLK, iq) for _ in range(max(1, args.steps))]
        tot = float(np.sum(secs))
        v = threads * blocks * BLK * len(secs) / tot / 1e6
        res = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "reference",
               "sample": "%d steps, each %d threads x %d blocks of %d IQ samples (%s), the reference's own classes compiled from "
                         "/root/reference with a generic-C VOLK shim, unthrottled" % (len(secs), threads, blocks, BLK, wl),
               "seconds": tot}
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
                "steps": len(secs), "warmup": max(1, args.warmup), "ms_per_step": 1000.0 * tot / len(secs),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": res,
                "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev_index = local_rank if world > 1 else 0
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    affinity = None
    try:  # run this rank (and first-touch its pinned staging buffers) on the CPUs next to its GPU
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(dev_index)
        pynvml.nvmlDeviceSetCpuAffinity(hnd)
        affinity = len(os.sched_getaffinity(0))
    except Exception:
        affinity = None

    dec = make_decoder(wl, C, T, nblk, dev_index)
    impl_desc = dec.describe() if hasattr(dec, "describe") else None  # which kernels this handle selected
    iq = gen_iq_device(torch, dev, fs, C, T, mode)
    width = 2 if (mode == "fm" and stereo) else 1
    audio_cap = int(T * 48000.0 / fs) * width + 64
    audio = torch.zeros((C, audio_cap), dtype=torch.float64, device=dev)
    bl = [BLK] * nblk
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def step():
        return dec.process_device(iq.data_ptr(), T, bl, audio.data_ptr(), audio_cap, sh)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(dev_index)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        lens = step()
    e1.record(stream)
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end)
    ms = e0.elapsed_time(e1)
    launches = dec.last_launches() * args.steps
    if dist is not None:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    total_samples = world * C * T * args.steps
    value = total_samples / (ms * 1e-3) / 1e6

    # per-stage device times of one step (CUDA events inside the library, same stream)
    dec.set_profiling(True)
    stage = {}
    reps = 3
    for _ in range(reps):
        step()
        torch.cuda.synchronize()
        for k, v in dec.stage_times().items():
            stage[k] = stage.get(k, 0.0) + v / reps
    dec.set_profiling(False)
    peak, peak_src = peaks()
    n_audio = int(lens.sum())  # audio values per channel per step (interleaved L, R when stereo)
    alg_bytes = C * T * 8 + C * n_audio * 8  # IQ read + audio written, per step
    roof = rooflines(wl, dec, stage, C, T, n_audio, width, ms / args.steps, peak, peak_src, fs)
    # opt-in: the same channels as G independent handles on G streams. Channels are independent, so this is only a
    # different schedule: the HBM-bound, shared-memory-bound and latency-bound kernels of different handles overlap.
    multi = None
    if args.handles > 1:
        G = args.handles
        Cg = C // G
        dec.close()  # its rings are not needed any more; G handles of C/G channels take their place
        decs = [make_decoder(wl, Cg, T, nblk, dev_index) for _ in range(G)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
        row = T * 8  # bytes per channel of the IQ buffer; audio rows are audio_cap doubles

        def mstep():
            for g in range(G):
                decs[g].process_device(iq.data_ptr() + g * Cg * row, T, bl, audio.data_ptr() + g * Cg * audio_cap * 8,
                                       audio_cap, streams[g].cuda_stream)

        for _ in range(max(3, args.warmup)):
            mstep()
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for sg in streams:
            sg.wait_event(m0)
        for _ in range(args.steps):
            mstep()
        for sg in streams:
            ev = torch.cuda.Event()
            ev.record(sg)
            stream.wait_event(ev)
        m1.record(stream)
        barrier()
        mms = m0.elapsed_time(m1)
        if dist is not None:
            tm = torch.tensor([mms], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            mms = float(tm.item())
        multi = {"handles": G, "channels_per_handle": Cg, "value": world * G * Cg * T * args.steps / (mms * 1e-3) / 1e6,
                 "unit": "Msamples/s", "ms_per_step": mms / args.steps,
                 "launches_per_step": sum(d.last_launches() for d in decs)}
        for d in decs:
            d.close()

    # end to end through the host-buffer entry point (pinned host memory, H2D + D2H inside)
    e2e = None
    if not args.no_e2e:
        dec2 = make_decoder(wl, Ce, T, nblk, dev_index)
        h_iq = torch.empty((Ce, T), dtype=torch.complex64, pin_memory=True)
        h_iq.copy_(iq[:Ce])
        h_np = h_iq.numpy()
        h_out = torch.empty((Ce, audio_cap), dtype=torch.float64, pin_memory=True).numpy()
        for _ in range(3):
            a, l = dec2.process_blocks(h_np, bl, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            a, l = dec2.process_blocks(h_np, bl, out=h_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": world * Ce * T * args.e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": Ce * T * 8, "d2h_bytes_per_step": int(a.shape[1]) * 8 * Ce,
               "channels": Ce, "note": "fmr_fm_process_host: pinned host IQ -> device -> audio back to host"}
        # The ceiling of this number on this box: the same pinned buffer copied to the device by every rank at the same
        # time, nothing else (PCIe + host memory; on the 8-GPU node all GPUs hang off one NUMA node).
        d_sink = torch.empty((Ce, T), dtype=torch.complex64, device=dev)
        d_sink.copy_(h_iq, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            d_sink.copy_(h_iq, non_blocking=True)
        torch.cuda.synchronize()
        dtc = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dtc], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dtc = float(tt.item())
        del d_sink
        e2e["h2d_ceiling"] = {"gb_per_s": world * Ce * T * 8 * args.e2e_steps / dtc / 1e9,
                              "value": world * Ce * T * args.e2e_steps / dtc / 1e6, "unit": "Msamples/s",
                              "note": "plain pinned-host -> device copy of the same cf32 buffer on all %d ranks at once" % world}
        e2e["frac_of_h2d_ceiling"] = e2e["value"] / e2e["h2d_ceiling"]["value"]
        if mode == "fm" and fs != 384000.0:
            # same call with the IQ as int16 pairs (16-bit WAV as FileSource reads it): half the PCIe bytes
            h_i16 = torch.empty((Ce, T, 2), dtype=torch.int16, pin_memory=True)
            h_i16.copy_((torch.view_as_real(iq[:Ce]) * 32768.0).clamp_(-32768, 32767).round_().to(torch.int16))
            q_np = h_i16.numpy()
            for _ in range(2):
                a, l = dec2.process_blocks_i16(q_np, bl, out=h_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                a, l = dec2.process_blocks_i16(q_np, bl, out=h_out)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            e2e["int16_ingest"] = {"value": world * Ce * T * args.e2e_steps / dt / 1e6, "unit": "Msamples/s",
                                   "h2d_bytes_per_step": Ce * T * 4,
                                   "note": "fmr_fm_process_host_i16: IQ as int16 pairs, converted in the first kernel"}
            h_i16 = q_np = None  # release the pinned staging before the next sub-measurement
        try:
            # the whole file path (SURVEY.md 8 f1 + f4) on 8-bit IQ, the narrowest format FileSource accepts
            # (format=U8_LE): sample decode, decoder, level metering, squelch gain and the int16 sink format all on
            # the device, so 2 B per IQ sample go up and 2 B per audio value come back
            from airspy_fmradion_b200 import _capi
            h_u8 = torch.empty((Ce, T, 2), dtype=torch.uint8, pin_memory=True)
            h_u8.copy_((torch.view_as_real(iq[:Ce]) * 127.0).round_().clamp_(-128, 127).add_(128).to(torch.uint8))
            u_np = h_u8.numpy().reshape(Ce, -1)
            o_i16 = torch.empty((Ce, audio_cap), dtype=torch.int16, pin_memory=True).numpy()
            for _ in range(2):
                a, l = dec2.process_blocks_io(u_np, _capi.IQ_U8, bl, out_format=_capi.OUT_S16, out=o_i16)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                a, l = dec2.process_blocks_io(u_np, _capi.IQ_U8, bl, out_format=_capi.OUT_S16, out=o_i16)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            e2e["u8_file_path"] = {"value": world * Ce * T * args.e2e_steps / dt / 1e6, "unit": "Msamples/s",
                                   "h2d_bytes_per_step": Ce * T * 2, "d2h_bytes_per_step": int(a.shape[1]) * 2 * Ce,
                                   "note": "fmr_fm_process_host_io: U8 IQ in, output stage (levels, -6 dB, int16) on "
                                           "the device, int16 audio out"}
        except Exception as ex:  # optional sub-measurement: the bench line must still print
            e2e["u8_file_path"] = {"error": str(ex)}
        dec2.close()

    single_homed = None
    if args.single_homed > 0 and dist is not None and mode == "fm":
        # every rank takes part in the same collectives; any failure here fails the run (opt-in measurement)
        from airspy_fmradion_b200 import _capi
        from airspy_fmradion_b200.shard import ShardedDecoder
        Cs = args.single_homed * world
        sd = ShardedDecoder(Cs, lambda n: make_decoder(wl, n, T, nblk, dev_index))
        raw_root = None
        if rank == 0:
            rows = torch.arange(Cs, device=dev) % C
            raw_root = (torch.view_as_real(iq[rows]) * 32768.0).clamp_(-32768, 32767).round_().to(torch.int16)
            raw_root = raw_root.reshape(Cs, T * 2).view(torch.uint8)
        for _ in range(2):
            sd.process_blocks_from_root(raw_root, _capi.IQ_S16, bl, _capi.OUT_S16, device=dev)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.e2e_steps):
            full, _l = sd.process_blocks_from_root(raw_root, _capi.IQ_S16, bl, _capi.OUT_S16, device=dev)
        s1.record(stream)
        barrier()
        tt = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        single_homed = {"value": Cs * T * args.e2e_steps / (float(tt.item()) * 1e-3) / 1e6, "unit": "Msamples/s",
                        "channels_total": Cs, "scatter_bytes_per_step": Cs * T * 4,
                        "note": "int16 IQ of all channels resident on rank 0 -> NCCL scatter -> decode + output stage on "
                                "the owning GPU -> NCCL gather of int16 audio to rank 0 (SURVEY.md 8 e)"}
        sd.dec.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))  # the CPU baseline gets every host core
        except Exception:
            pass
        try:
            cpu = cpu_reference(wl, args.cpu_seconds, os.cpu_count() or 1)
        except Exception as ex:  # the bench line must still print
            cpu = {"error": str(ex)}

    if rank == 0:
        # what the path computes in: linear filters FP32 (the FM audio filters FP64 with FMR_AUDIO_FP64=1), recurrences FP64
        if mode == "fm" and os.environ.get("FMR_AUDIO_FP64", "0") not in ("", "0"):
            dtype = "f32 (IF resampler) / f64 (PLL, deemphasis, audio filters, DC block)"
        elif mode == "fm":
            dtype = "f32 (IF resampler, audio filters) / f64 (PLL, deemphasis, DC block)"
        else:
            dtype = "f32 (IF resampler, channel filter) / f64 (audio)"
        line = {"metric": metric, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": dtype,
                "data": "synthetic",
                "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        if impl_desc:
            line["implementation"] = impl_desc
        if isinstance(roof, dict) and "fp32" in roof:
            line["roofline_fp32"] = roof.pop("fp32")
        if single_homed is not None:
            line["single_homed"] = single_homed
        if multi is not None:
            line["multi_handle"] = multi
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
