#!/bin/bash
# First GPU call of the next round: parity + A/B of the kernel variants that were written and host-checked at the
# end of round 1 but have not been measured (the round's GPU budget was spent).
#   FMR_FFT_INPLACE=2  k_fir_fft_ip32 (radix 32 x 32 x 16 in-place form of the 16384-point FFT low-pass)
# Flip the default in fmr_host.cuh (fft_inplace32) only if the parity tests pass and the sweep shows a gain.
set -u
mkdir -p gpurun_out
#   FMR_FFT_INPLACE8K=1 k_fir_fft_ip8k (in-place form of the 8192-point blocks: remainder block at 10 Msps, all blocks at 1 Msps)
FMR_FFT_INPLACE=2 FMR_FFT_INPLACE8K=1 timeout 200 python -m pytest tests/test_fm_gpu.py tests/test_golden_gpu.py tests/test_edge_gpu.py -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_ip32.log
#   FMR_FFT_EPI=1      k_fir_fft_ip<512,512,1>: polyphase bank + window offsets in shared memory (the epilogue holds a third of
#                      the FFT kernel's stall samples); results must be bit-identical (same checksum as all_on_329)
FMR_FFT_EPI=1 timeout 120 python -m pytest tests/test_fm_gpu.py -q -x -k "if_stage or cfg2" 2>&1 | tail -3 | tee gpurun_out/pytest_epi.log
timeout 200 python tools/sweep_variants.py all_on_329 fft_epi_smem fft_inplace_r32 fft_inplace_r32_epi fft_inplace_8k fft_inplace_r32_8k fft_stockham 2>gpurun_out/sweep_ip32.err | tee gpurun_out/sweep_ip32.log
# Independent handles per GPU on their own streams (bench.py --handles): overlap of the HBM-bound half-band stream,
# the shared-memory-bound FFT and the latency-bound core ACROSS handles, with no library change. The second / third
# runs make the kernels small enough to share an SM: FFT capped at 72 registers (FMR_FFT_REGCAP=1), half-band stream
# with two TMA stages = 68 KB (FMR_HBS_TMA=5, measured as fast as three stages).
# Rate pairs added without GPU budget: if green, set their `verified` flag to 1 in tools/gen_tables.py, regenerate, and
# drop the gate of tests/test_newrates_gpu.py.
FMR_EXPERIMENTAL_RATES=1 timeout 300 python -m pytest tests/test_newrates_gpu.py -q -s 2>&1 | tail -20 | tee gpurun_out/pytest_newrates.log
B="python bench.py --no-cpu --no-e2e --steps 6 --warmup 3"
for g in 2 4; do
  $B --handles $g 2>/dev/null | tail -1 > gpurun_out/multi_handle_${g}.json
  FMR_FFT_REGCAP=1 FMR_HBS_TMA=5 $B --handles $g 2>/dev/null | tail -1 > gpurun_out/multi_handle_${g}_coresident.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/multi_handle_*.json")):
    d = json.load(open(f))
    print(f, "single %.1f" % (d["value"] / 1e3), "multi", d.get("multi_handle"))
PY
