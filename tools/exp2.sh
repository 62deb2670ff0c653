#!/bin/bash
B="python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --channels 4096"
run() { $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print('$1', round(d['value']), round(d['ms_per_step'],2), s['if_halfband_cascade'], s['audio_halfband_cascade'], s['fm_pll_stereo_deemph'])"; }
python -m pytest tests -q -m gpu 2>&1 | tail -3
run "default(256,128)"
for v in "128 64" "256 64" "128 128"; do set -- $v
  FMR_NVCC_EXTRA="-DFMR_HB_TILE=$1 -DFMR_HB_THREADS=$2" python __graft_entry__.py > /dev/null 2>&1
  run "tile=$1,thr=$2"
done
FMR_NVCC_EXTRA="-DFMR_HB_TILE=256" python __graft_entry__.py > /dev/null 2>&1
