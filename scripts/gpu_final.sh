#!/bin/bash
# Round-end validation in one gpurun call: the whole GPU suite in one process (the way the driver runs it), the bench line with either optimiser, the extra workloads (configs[2], configs[4] sweep), and an
# ncu capture of the optimiser kernels.  Every piece runs under its own timeout.
mkdir -p gpurun_out
T0=$SECONDS
echo "=== full suite"; timeout 400 python -m pytest tests -q -m gpu -rf --timeout 180 -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "exit $?"; tail -25 gpurun_out/final_tests.log | cut -c1-400
echo "t=$((SECONDS-T0))"
echo "=== bench (torch optimiser)"; timeout 300 python bench.py --optim torch > gpurun_out/bench_torch.json 2> gpurun_out/bench_torch.err; echo "exit $?"; cut -c1-700 gpurun_out/bench_torch.json; tail -3 gpurun_out/bench_torch.err
echo "t=$((SECONDS-T0))"
echo "=== bench (fused optimiser)"; timeout 300 python bench.py --optim fused --no-decode --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "exit $?"; cut -c1-700 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
echo "t=$((SECONDS-T0))"
for w in cfg3 discrete_token continuous_token; do
  echo "=== bench $w"; timeout 200 python bench.py --workload $w --steps 6 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit $?"; cut -c1-330 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err
  echo "t=$((SECONDS-T0))"
done
echo "=== optimiser micro"; timeout 120 python scripts/optim_micro.py 2>&1 | tail -5
echo "=== ncu optimiser kernels"
ITERS=1 ONLY=fused timeout 150 ncu --set full --clock-control none --import-source on -k regex:"grad_sqnorm_kernel|adam_update_kernel|adam_prepare_kernel" -s 9 -c 9 -f -o gpurun_out/ncu_optim python scripts/optim_micro.py > gpurun_out/ncu_optim.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_optim.log
echo "t=$((SECONDS-T0))"
