#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_fm_gpu.py tests/test_golden_gpu.py tests/test_edge_gpu.py -m gpu -x -q -s 2>&1 ) | grep -E "cfg3|E8|passed|failed|Error|assert" | head
python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --blocks 128 --workload cfg3_fm_stereo_10Msps_E200 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
