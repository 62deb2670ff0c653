#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --channels 1024"
ncu --set full --clock-control none --import-source on -k regex:k_fir_fft -s 4 -c 1 -f -o gpurun_out/prof_fft16_r01c $B > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
