#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 --blocks 128 --workload cfg3_fm_stereo_10Msps_E200"
ncu --set full --clock-control none --import-source on -k regex:k_mpf -s 3 -c 1 -f -o gpurun_out/prof_mpf_r01 $B > /dev/null 2>&1
ls -la gpurun_out/prof_mpf_r01.ncu-rep
