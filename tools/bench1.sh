#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 724 -c 370 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 740 -c 6 -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?"
