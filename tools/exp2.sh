#!/bin/bash
# sub-batch x streams x PDL sweep (GPU box)
run() { env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-zopt > /tmp/o.json 2> /tmp/e.txt || tail -3 /tmp/e.txt; python -c "
import json
d=json.loads(open('/tmp/o.json').read()); print('$*', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
run ESR_SUBBATCH=16 ESR_STREAMS=1
run ESR_SUBBATCH=8 ESR_STREAMS=2
run ESR_SUBBATCH=8 ESR_STREAMS=2 ESR_NO_PDL=1
run ESR_SUBBATCH=4 ESR_STREAMS=2
run ESR_SUBBATCH=4 ESR_STREAMS=2 ESR_NO_PDL=1
run ESR_SUBBATCH=4 ESR_STREAMS=4 ESR_NO_PDL=1
run ESR_SUBBATCH=2 ESR_STREAMS=2 ESR_NO_PDL=1
run ESR_SUBBATCH=2 ESR_STREAMS=4 ESR_NO_PDL=1
