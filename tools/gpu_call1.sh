#!/bin/bash
# gpurun call: GPU test suite, then the variant sweep.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1100 python -m pytest tests -m gpu -x -q -s 2>&1 ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep_variants.py > gpurun_out/sweep_variants.log 2> gpurun_out/sweep_variants.err
cat gpurun_out/sweep_variants.log | cut -c1-400
tail -3 gpurun_out/sweep_variants.err
