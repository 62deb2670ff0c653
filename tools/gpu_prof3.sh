#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --channels 1024"
ncu --set full --clock-control none --import-source on -k regex:k_fir_fft -s 4 -c 3 -f -o gpurun_out/prof_fft_r01 $B > gpurun_out/prof_fft.log 2>&1
tail -3 gpurun_out/prof_fft.log
ls -la gpurun_out/prof_fft_r01.ncu-rep
