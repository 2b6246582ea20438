#!/bin/bash
# First-contact probe on the GPU box: every test file in its own process (a trapped kernel poisons its CUDA context),
# each under its own timeout, logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in test_gpu_kernels test_gpu_gemm test_gpu_rank test_gpu_encoder; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -x -s --tb=short > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
