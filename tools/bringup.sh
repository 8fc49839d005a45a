#!/bin/bash
# GPU bring-up: each stage in its own process so a device-side trap does not poison later stages.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run cem python -m pytest tests/test_gpu_cem.py -q -m gpu -p no:cacheprovider
run conv_simt python -m pytest tests/test_gpu_conv.py -q -m gpu -p no:cacheprovider -k simt
run conv_tc python -m pytest tests/test_gpu_conv.py -q -m gpu -p no:cacheprovider -k "not simt"
run net_simt python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -k "simt or identical or fallback"
run net_tc python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -k "tc or consistency"
