#!/bin/bash
for v in 16 8 4 2; do
  ESR_SUBBATCH=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sub=$v', 'ms', round(d['ms_per_step'],2), d['roofline']['kernel'][:70], d['phases_ms'], d['clocks']['sm_mhz'])"
done
