#!/bin/bash
B="python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --channels 4096"
run() { $B $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],2))"; }
run "blocks220,default" "--blocks 220"
for sms in 16 24 32; do
FMR_TIME_CHUNKS=4 FMR_CHUNK_MIN_BLOCKS=55 FMR_SERIAL_SMS=$sms run "blocks220,chunks4x55,sms$sms" "--blocks 220"
done
FMR_TIME_CHUNKS=8 FMR_CHUNK_MIN_BLOCKS=55 FMR_SERIAL_SMS=24 run "blocks440,chunks8x55,sms24" "--blocks 440"
FMR_TIME_CHUNKS=8 FMR_CHUNK_MIN_BLOCKS=27 FMR_SERIAL_SMS=24 run "blocks220,chunks8x27,sms24" "--blocks 220"
FMR_TRACE=1 FMR_TIME_CHUNKS=4 FMR_CHUNK_MIN_BLOCKS=55 FMR_SERIAL_SMS=24 $B --steps 1 --blocks 220 2>&1 | grep "fmr" | grep -v "chunk 0 .* 0.00[0-9] ->" | sed -n 2,18p
