timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final3_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/final3_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final3_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/final3_smoke.log
timeout 400 python bench.py > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/final3_bench.json
