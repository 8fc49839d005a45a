#!/bin/bash
# fused-kernel mechanism costs: rebuild variants on the GPU box (timing experiments; results of these variants are invalid)
b() { python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-zopt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'],2), d['phases_ms']['convs'], d['clocks']['sm_mhz'])"; }
for v in "-DESR_RDB_NO_DEPS -DESR_RDB_NO_SIGNAL -DESR_RDB_NO_PUBLISHER" "-DESR_RDB_NO_DEPS -DESR_RDB_NO_SIGNAL -DESR_RDB_NO_PUBLISHER -DESR_RDB_NO_POLL"; do
  ESR_NVCC_EXTRA="$v" python explorable-super-resolution_old_b200/build.py --force > /dev/null 2>&1 || echo build failed
  ESR_RDB_LAYERS=1 ESR_RDB_CHUNK=16 b "[$v] layers=1 chunk=16"
done
ESR_FUSE_RDB=0 b unfused
