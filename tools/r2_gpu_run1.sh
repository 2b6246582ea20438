#!/bin/bash
# Round-2 first contact: new attention kernel (correctness vs the first-generation kernel, timing of the variants), full
# GPU test suite, bench (with the 100k webvid leg), library baseline.  Every step under its own timeout; logs in gpurun_out/.
OUT=gpurun_out/r2_run1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi.txt 2>&1
# 1. attention: tc2 vs tc1 numerics, then timings
timeout 300 python tools/attention_check.py > $OUT/attention_check.log 2>&1; echo "attention_check exit $?" | tee -a $OUT/summary.txt
for v in "" poly0 poly8 timing; do
  timeout 120 python tools/attention_bench.py $v >> $OUT/attention_bench.log 2>&1; echo "attention_bench '$v' exit $?" | tee -a $OUT/summary.txt
done
FC_ATTENTION=tc1 timeout 120 python tools/attention_bench.py >> $OUT/attention_bench_tc1.log 2>&1
FC_ATTENTION=tc1 timeout 120 python tools/attention_bench.py timing >> $OUT/attention_bench_tc1.log 2>&1
# 2. tests
timeout 1500 python -m pytest tests -m gpu -q -x -s --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log
# 3. bench (N=1, default flags), then the library baseline
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
timeout 600 python tools/torch_gpu_baseline.py --compile > $OUT/torch_gpu_baseline.json 2> $OUT/torch_gpu_baseline.err; echo "torch baseline exit $?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
