#!/bin/bash
# ncu --set full captures at the bench's pass size (500 videos x 4 frames = one 2000-frame pass).  GEMM launch order of a
# vision pass: patch-embed, then per block QKV (folded ln_1), out-proj + residual, fc1 (folded ln_2 + QuickGELU), fc2 + residual.
OUT=gpurun_out/r2_ncu
mkdir -p $OUT
timeout 600 ncu --set full --import-source on --clock-control none -k gemm_bf16_tn_kernel -s 5 -c 4 -o $OUT/gemm_block1 python tools/profile_step.py 1 500 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm exit $?" | tee -a $OUT/summary.txt
ls -la $OUT
