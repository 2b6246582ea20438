#!/bin/bash
OUT=gpurun_out/r2_run9
mkdir -p $OUT
timeout 300 python tools/long_attention_check.py > $OUT/long_default.log 2>&1; echo "long default exit $?" | tee -a $OUT/summary.txt
FITCLIP_VARIANT=kb64 timeout 300 python tools/long_attention_check.py > $OUT/long_kb64.log 2>&1; echo "long kb64 exit $?" | tee -a $OUT/summary.txt
cat $OUT/long_default.log $OUT/long_kb64.log
FITCLIP_VARIANT=kb64 timeout 600 python tools/geometry_bench.py clip_vit_l_14 clip_vit_l_14_336px > $OUT/geometry_kb64.jsonl 2> $OUT/geometry_kb64.err; echo "geometry kb64 exit $?" | tee -a $OUT/summary.txt
cat $OUT/geometry_kb64.jsonl
cat $OUT/summary.txt
