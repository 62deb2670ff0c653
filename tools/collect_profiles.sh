#!/bin/bash
# Run on the GPU box (gpurun): the parity suite, the ncu evidence and the bench lines that profiles/ summarises
# (tools/summarize_profiles.py turns gpurun_out/ into the tracked files).   bash tools/collect_profiles.sh r02
set -u
mkdir -p gpurun_out
R=${1:-r02}
B="python bench.py --no-cpu --no-e2e"
# 0. parity suite and the smoke entry point
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_${R}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${R}.log 2>&1
# 1. every launch of the default bench with its device time (serialised, cold cache: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file gpurun_out/launches_${R}.csv $B --steps 2 --warmup 3 > gpurun_out/bench_under_ncu_${R}.log 2>&1
# 2. full captures, one launch each, 1024 channels (-s counts matching launches only)
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1_${R} \
          $B --steps 1 --warmup 2 --channels 1024 > /dev/null 2>&1; }
cap frontend_fused k_frontend_fused 2
cap fdr k_fdr 2
cap core k_fm_core_fused 2
cap audio_hb "k_hb_cascade<float, 2, 0, 7" 2
cap pilot_cut "k_fir_quirk<float" 2
cap tail k_fm_tail 2
# 3. device-resident value over channel counts (multiples of 148 SMs x 32 channels, and the round-1 default) and workloads
for ch in 148 1184 4736 9472 14208 16384 18944; do
  timeout 400 $B --steps 4 --warmup 3 --channels $ch 2>&1 | tail -1 > gpurun_out/sweep_cfg2_${ch}_${R}.json
done
timeout 400 $B --steps 3 --warmup 3 --blocks 128 --workload cfg3_fm_stereo_10Msps_E200 2>&1 | tail -1 > gpurun_out/sweep_cfg3_${R}.json
timeout 400 $B --steps 4 --warmup 3 --blocks 128 --workload cfg4_fm_stereo_1Msps 2>&1 | tail -1 > gpurun_out/sweep_cfg4_${R}.json
timeout 400 $B --steps 4 --warmup 3 --blocks 128 --workload cfg5_am_384ksps 2>&1 | tail -1 > gpurun_out/sweep_cfg5_${R}.json
# 4. the default bench line (with e2e and the CPU reference on this box's cores) and the reference arm
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default_${R}.json
timeout 900 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_reference_${R}.json
# 5. the all-FP64 audio chain (FMR_AUDIO_FP64=1) for comparison
FMR_AUDIO_FP64=1 timeout 400 $B --steps 6 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_audio_fp64_${R}.json
ls -la gpurun_out | head -60
