#!/bin/bash
# Round profile capture (GPU box): bench line, per-launch list with DRAM bytes, one full-set capture of the top kernel.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-zopt > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 724 -c 370 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-zopt > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 740 -c 5 -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-zopt > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
