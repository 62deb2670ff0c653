#!/bin/bash
# experiment: half-band cascade tile/thread variants (rebuilds the library on the GPU box)
B="python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --channels 4096"
run() { $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],2), d['roofline']['stage_ms']['if_halfband_cascade'], d['roofline']['stage_ms']['audio_halfband_cascade'])"; }
python -m pytest tests -q -m gpu 2>&1 | tail -3
run "default(512,256)"
for v in "256 128" "256 256" "1024 256"; do set -- $v
  FMR_NVCC_EXTRA="-DFMR_HB_TILE=$1 -DFMR_HB_THREADS=$2" python __graft_entry__.py > /dev/null 2>&1
  run "tile=$1,thr=$2"
done
python __graft_entry__.py > /dev/null 2>&1  # note: _newer() sees a fresh lib; force default rebuild below
FMR_NVCC_EXTRA="-DFMR_HB_TILE=256" python __graft_entry__.py > /dev/null 2>&1
