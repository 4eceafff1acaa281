#!/bin/bash
# Runs the GPU test groups in separate processes (a device trap in one group must not poison the next).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 python -m pytest "$@" -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run simt tests/test_gpu_kernels.py -k "not tcgen05 and not epilogues and not split_k and not rejects and not tensor_core"
run tcgen05 tests/test_gpu_kernels.py -k "tcgen05 or epilogues or split_k or rejects"
run attn_tc tests/test_gpu_kernels.py -k "tensor_core"
run model_fp32 tests/test_gpu_model.py -k "not bf16"
run model_bf16 tests/test_gpu_model.py -k "bf16"
run decode tests/test_gpu_decode.py
run fullsize tests/test_gpu_fullsize.py
run sampling_loss tests/test_gpu_sampling.py tests/test_gpu_loss.py
run optimizer tests/test_gpu_optimizer.py
run regression tests/test_gpu_regression.py
run tokens tests/test_gpu_tokens.py
run parity_oracle tests/test_gpu_parity_oracle.py
