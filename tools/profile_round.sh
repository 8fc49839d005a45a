#!/bin/bash
# Round profile capture (GPU box): bench line, then the 351 conv launches of one warm step with time, tensor-pipe activity and DRAM bytes.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-zopt > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --cache-control none -k regex:conv3x3 -s 1053 -c 351 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-zopt > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
