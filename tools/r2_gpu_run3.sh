#!/bin/bash
OUT=gpurun_out/r2_run3
mkdir -p $OUT
timeout 300 python tools/attention_check.py > $OUT/attention_check.log 2>&1; echo "attention_check exit $?" | tee -a $OUT/summary.txt
tail -12 $OUT/attention_check.log
for v in "" poly0 poly8 timing; do
  timeout 120 python tools/attention_bench.py $v >> $OUT/attention_bench.log 2>&1; echo "attention_bench '$v' exit $?" | tee -a $OUT/summary.txt
done
FC_ATTENTION=tc2 timeout 120 python tools/attention_bench.py >> $OUT/attention_bench_tc2.log 2>&1
cat $OUT/attention_bench.log
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log
timeout 900 python bench.py --webvid-videos 0 > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_tc3 -s 5 -c 1 -o $OUT/att_tc3 python tools/attention_bench.py > $OUT/ncu_att.log 2>&1; echo "ncu exit $?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
