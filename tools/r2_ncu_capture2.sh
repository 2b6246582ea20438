#!/bin/bash
# ncu --set full captures of the kernels added late in round 2 (see tools/ncu_targets.py); summaries -> profiles/ via tools/ncu_summary.py
OUT=gpurun_out/r2_ncu2
mkdir -p $OUT
timeout 300 python tools/ncu_targets.py > $OUT/plain.log 2>&1; echo "plain run exit $?" | tee -a $OUT/summary.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'attention_tc_long_kernel|gemm_bf16_tn_kernel' -o $OUT/late_kernels python tools/ncu_targets.py > $OUT/ncu.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
ls -la $OUT
