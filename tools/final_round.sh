#!/bin/bash
# Run on the GPU box (gpurun) at the end of a round: the GPU suite, smoke(), the two bench arms, then the ncu
# evidence that changed this round. Most important first; every step has its own time limit.
set -u
mkdir -p gpurun_out
R=${1:-r01}
B="python bench.py --no-cpu --no-e2e"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_${R}.log
timeout 400 python bench.py 2>gpurun_out/bench_default_${R}.err | tail -1 > gpurun_out/bench_default_${R}.json; head -c 400 gpurun_out/bench_default_${R}.json; echo
timeout 200 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_reference_${R}.json; head -c 300 gpurun_out/bench_reference_${R}.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file gpurun_out/launches_${R}.csv $B --steps 2 --warmup 3 --channels 1024 > gpurun_out/bench_under_ncu_${R}.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fir_fft -s 4 -c 3 -f -o gpurun_out/prof_fft_${R} $B --steps 1 --warmup 3 --channels 1024 > /dev/null 2>&1
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_io_gpu.py -q -x -k "output_stage or am_nbfm or capacity" 2>&1 | tail -4 | tee gpurun_out/sanitizer_io_${R}.log
ls -la gpurun_out | tail -12
