#!/bin/bash
# smoke + short bench + ncu launch list (+ optional full capture of the GEMM kernel)
mkdir -p gpurun_out
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py --steps ${STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$NCU_LIST" ]; then
echo "=== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-600} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
tail -3 gpurun_out/ncu_bench.log
fi
if [ -n "$NCU_FULL" ]; then
echo "=== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${NCU_FULL} -s ${NCU_SKIP:-30} -c 3 -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1; echo "exit $?"
tail -3 gpurun_out/ncu_full.log
fi
