#!/bin/bash
# One GPU session (run under gpurun): parity suite, A/B sweep of the named variants, optional ncu capture.
#   bash tools/gpu_session.sh "<pytest args>" "<sweep variants>" "<ncu kernel regex>" <tag>
set -u
mkdir -p gpurun_out
PYT=${1:-}; SWEEP=${2:-}; NCU=${3:-}; TAG=${4:-r02}
if [ -n "$PYT" ]; then timeout 900 python -m pytest $PYT -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.log; fi
if [ -n "$SWEEP" ]; then timeout 600 python tools/sweep_variants.py $SWEEP 2>gpurun_out/sweep_${TAG}.err | tee gpurun_out/sweep_${TAG}.jsonl; fi
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$NCU -s 2 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --no-cpu --no-e2e --steps 1 --warmup 2 --channels 1024 > gpurun_out/ncu_${TAG}.log 2>&1
  ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
fi
