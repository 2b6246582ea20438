#!/bin/bash
OUT=gpurun_out/r2_run7
mkdir -p $OUT
timeout 300 python tools/attention_check.py > $OUT/attention_check.log 2>&1; echo "attention_check exit $?" | tee -a $OUT/summary.txt
tail -4 $OUT/attention_check.log
timeout 120 python tools/attention_bench.py >> $OUT/attention_bench.log 2>&1; echo "attention_bench exit $?" | tee -a $OUT/summary.txt
timeout 120 python tools/attention_bench.py timing >> $OUT/attention_bench.log 2>&1
FC_ATTENTION=tc2a timeout 120 python tools/attention_bench.py >> $OUT/attention_bench_tc2a.log 2>&1
FC_ATTENTION=tc2a timeout 120 python tools/attention_bench.py timing >> $OUT/attention_bench_tc2a.log 2>&1
cat $OUT/attention_bench.log $OUT/attention_bench_tc2a.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu.log
timeout 600 python tools/geometry_bench.py clip_vit_l_14 clip_vit_l_14_336px > $OUT/geometry_bench.jsonl 2> $OUT/geometry_bench.err; echo "geometry exit $?" | tee -a $OUT/summary.txt
cat $OUT/geometry_bench.jsonl
timeout 900 python bench.py --webvid-videos 0 --train-videos 0 > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
python -c "
import json; d=json.load(open('$OUT/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e_roofline_frac'], d['e2e']['value'], d['extra']['e2e_uint8']['value'], d['clocks']); print(d['roofline']['ms_by_kernel_class'], d['roofline']['achieved'])"
cat $OUT/summary.txt
