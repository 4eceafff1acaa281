#!/bin/bash
# Second round-end call: the fixed optimiser test, the final default bench line, and the ncu launch list of the same
# command (steady-state step shares).
mkdir -p gpurun_out
T0=$SECONDS
echo "=== optimiser tests"; timeout 200 python -m pytest tests/test_gpu_optimizer.py -q -m gpu --timeout 120 -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
echo "t=$((SECONDS-T0))"
echo "=== bench (default)"; timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "exit $?"; cut -c1-500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "t=$((SECONDS-T0))"
echo "=== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/ncu_bench_final.log 2>&1; echo "exit $?"
tail -2 gpurun_out/ncu_bench_final.log | cut -c1-300; wc -l gpurun_out/launches_final.csv
echo "t=$((SECONDS-T0))"
