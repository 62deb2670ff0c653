#!/bin/bash
# ncu captures of the current top kernels (C=1024 keeps the replays short)
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --channels 1024"
ncu --set full --clock-control none --import-source on -k regex:k_fm_core_fused -s 3 -c 1 -f -o gpurun_out/prof_core_r01b $B > /dev/null 2>&1
FMR_CORE_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:k_fm_pll2 -s 3 -c 1 -f -o gpurun_out/prof_pll2_r01b $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hb_stream -s 3 -c 1 -f -o gpurun_out/prof_hbs_r01b $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fir_fft -s 3 -c 1 -f -o gpurun_out/prof_fftf_r01b $B > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
