#!/bin/bash
# Round-end evidence in one gpurun call: the GPU test groups, the default bench line, the ncu launch list of one
# steady-state step and the ncu --set full captures of every kernel class (summarised by scripts/ncu_summary.py).
mkdir -p gpurun_out
bash scripts/gpu_check.sh 2>&1 | grep -E "^===|^exit|passed|failed|error"
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; tail -c 300 gpurun_out/bench_final.err
bash scripts/gpu_launch_list.sh > gpurun_out/launch_list.txt 2>&1; tail -40 gpurun_out/launch_list.txt
bash scripts/gpu_ncu_kernels.sh
for n in gemm attn elementwise decode; do python scripts/ncu_summary.py gpurun_out/ncu_$n.ncu-rep > gpurun_out/ncu_$n.txt 2>&1; done
ls -la gpurun_out | head -40
