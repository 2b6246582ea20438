#!/bin/bash
OUT=gpurun_out/r2_run5
mkdir -p $OUT
timeout 300 python tools/attention_check.py > $OUT/attention_check.log 2>&1; echo "attention_check exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/attention_check.log
for v in "" splits r120 splitpoly timing; do
  timeout 120 python tools/attention_bench.py $v >> $OUT/attention_bench.log 2>&1; echo "attention_bench '$v' exit $?" | tee -a $OUT/summary.txt
done
cat $OUT/attention_bench.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short -k attention > $OUT/pytest_attention.log 2>&1; echo "pytest attention exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_attention.log
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log
timeout 900 python tools/geometry_bench.py > $OUT/geometry_bench.jsonl 2> $OUT/geometry_bench.err; echo "geometry exit $?" | tee -a $OUT/summary.txt
cat $OUT/geometry_bench.jsonl
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
