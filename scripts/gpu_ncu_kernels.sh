#!/bin/bash
# ncu --set full captures of every kernel class on the small-footprint micro-benchmarks (fast replays),
# summarised under profiles/ by scripts/ncu_summary.py.
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
cap() { name=$1; regex=$2; skip=$3; count=$4; shift 4; timeout 400 $N -k regex:"$regex" -s $skip -c $count -f -o gpurun_out/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name exit $?"; }
ITERS=2 cap gemm "gemm_tc2_kernel" 5 12 python scripts/gemm_micro.py
ITERS=1 cap attn "attn_fwd2_kernel|attn_bwd_tc_kernel|attn_bwd_q_tc_kernel" 3 3 python scripts/attn_micro.py
ITERS=2 cap elementwise "add_ln_|colsum_kernel|embed_fwd_kernel" 3 12 python scripts/elementwise_micro.py
GRAPH=0 TS=1024 cap decode "attn_decode_stream_kernel|attn_decode_kernel" 60 2 python scripts/decode_micro.py
