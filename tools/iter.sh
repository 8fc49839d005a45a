#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -x -k "not simt" > gpurun_out/tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read()); print('ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['kernel'][:70], 'frac', round(d['roofline']['frac'],3), d['phases_ms'], d['clocks'])"; tail -5 gpurun_out/bench.err
python tools/conv5.py blocked
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 724 -c 370 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
