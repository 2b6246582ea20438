#!/bin/bash
OUT=gpurun_out/r2_run10
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_kernels.py -m gpu -q -x --tb=short > $OUT/pytest_part.log 2>&1; echo "pytest part exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_part.log
FC_ATTENTION=long timeout 120 python tools/attention_bench.py > $OUT/att_long_kb128.log 2>&1
FC_ATTENTION=long timeout 120 python tools/attention_bench.py kb64 > $OUT/att_long_kb64.log 2>&1
cat $OUT/att_long_kb128.log $OUT/att_long_kb64.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitizer_smoke.py > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/summary.txt
tail -6 $OUT/sanitizer_memcheck.log
cat $OUT/summary.txt
