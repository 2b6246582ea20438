#!/bin/bash
OUT=gpurun_out/r2_run8
mkdir -p $OUT
timeout 120 python tools/attention_bench.py >> $OUT/attention_bench.log 2>&1; cat $OUT/attention_bench.log
for f in 1000 2000 4000; do
  timeout 600 python bench.py --webvid-videos 0 --train-videos 0 --cpu-sample 8 --frames-per-pass $f > $OUT/bench_fpp$f.json 2> $OUT/bench_fpp$f.err; echo "bench fpp $f exit $?" | tee -a $OUT/summary.txt
  python -c "
import json; d=json.load(open('$OUT/bench_fpp$f.json')); print('fpp', $f, d['value'], d['ms_per_step'], d['e2e_roofline_frac'], d['e2e']['value'], d['extra']['e2e_uint8']['value'], d['clocks']['sm_mhz']); print(d['roofline']['ms_by_kernel_class'], d['roofline']['achieved'])"
done
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
python -c "
import json; d=json.load(open('$OUT/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e_roofline_frac'], d['e2e']['value'], d['extra']['e2e_uint8']['value'], d['clocks']); print(d['roofline']['ms_by_kernel_class'], d['roofline']['achieved']); w=d['extra']['webvid']; print({k:w[k] for k in ('seconds','videos_per_s','roofline_frac','sampled_rows_match_oracle','metrics')}); print(d['extra']['train_step'])"
cat $OUT/summary.txt
