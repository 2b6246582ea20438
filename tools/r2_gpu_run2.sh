#!/bin/bash
OUT=gpurun_out/r2_run2
mkdir -p $OUT
# full GPU suite (no -x: every failure is wanted)
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log
timeout 900 python tools/torch_gpu_baseline.py --compile > $OUT/torch_gpu_baseline.json 2> $OUT/torch_gpu_baseline.err; echo "torch baseline exit $?" | tee -a $OUT/summary.txt
# ncu source-level capture of the attention kernel (image shape only: -k filters the kernel, -c 1 one launch after warm-up)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_tc2 -s 5 -c 1 -o $OUT/att_tc2 python tools/attention_bench.py > $OUT/ncu_att.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
