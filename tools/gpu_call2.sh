#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_am_gpu.py tests/test_golden_gpu.py -m gpu -x -q -s 2>&1 ) | grep -E "AM |am_384k|passed|failed|Error|assert" | head -20
for ch in 64 8192; do
python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --blocks 128 --workload cfg5_am_384ksps --channels $ch 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['channels_per_gpu'], d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done
FMR_CORE_FUSED=0 python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --blocks 128 --workload cfg5_am_384ksps --channels 64 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unfused', d['config']['channels_per_gpu'], d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
