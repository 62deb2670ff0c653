#!/bin/bash
# Run on the GPU box (gpurun): collects the ncu evidence and the bench sweep that profiles/ summarises.
set -u
mkdir -p gpurun_out
R=${1:-r01}
B="python bench.py --no-cpu --no-e2e"
# 1. every launch of one bench run with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file gpurun_out/launches_${R}.csv $B --steps 2 --warmup 3 --channels 1024 > gpurun_out/bench_under_ncu_${R}.log 2>&1
# 2. full captures of the top kernels (one launch each); -s counts matching launches only
ncu --set full --clock-control none --import-source on -k regex:k_hb_stream_tma -s 3 -c 1 -f -o gpurun_out/prof_hbs_${R} $B --steps 1 --warmup 3 --channels 1024 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fir_fft -s 4 -c 3 -f -o gpurun_out/prof_fft_${R} $B --steps 1 --warmup 3 --channels 1024 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fm_core_fused -s 3 -c 1 -f -o gpurun_out/prof_core_${R} $B --steps 1 --warmup 3 --channels 1024 > /dev/null 2>&1
# 3. bench sweep (device-resident value) over channel counts and workloads
for ch in 148 1024 4096 8192 16384; do  # 16384 is the bench default
  $B --steps 4 --warmup 3 --channels $ch 2>&1 | tail -1 > gpurun_out/sweep_cfg2_${ch}_${R}.json
done
$B --steps 4 --warmup 3 --blocks 128 2>&1 | tail -1 > gpurun_out/sweep_cfg2_8192_b128_${R}.json
$B --steps 3 --warmup 3 --blocks 128 --workload cfg3_fm_stereo_10Msps_E200 2>&1 | tail -1 > gpurun_out/sweep_cfg3_${R}.json
$B --steps 4 --warmup 3 --blocks 128 --workload cfg4_fm_stereo_1Msps 2>&1 | tail -1 > gpurun_out/sweep_cfg4_${R}.json
$B --steps 4 --warmup 3 --blocks 128 --workload cfg5_am_384ksps 2>&1 | tail -1 > gpurun_out/sweep_cfg5_${R}.json
# 4. the default bench line (with e2e and the CPU reference on this box's cores) and the reference arm
python bench.py 2>&1 | tail -1 > gpurun_out/bench_default_${R}.json
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_reference_${R}.json
ls -la gpurun_out | head -40
