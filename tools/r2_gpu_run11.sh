#!/bin/bash
OUT=gpurun_out/r2_run11
mkdir -p $OUT
for v in "" spin; do
  FITCLIP_VARIANT=$v timeout 120 python tools/attention_bench.py >> $OUT/attention_bench.log 2>&1
done
cat $OUT/attention_bench.log
FITCLIP_VARIANT=gspin timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rank.py -m gpu -q -x --tb=short > $OUT/pytest_gspin.log 2>&1; echo "pytest gspin exit $?" | tee -a $OUT/summary.txt
tail -2 $OUT/pytest_gspin.log
for v in "" gspin allspin ""; do
  FITCLIP_VARIANT=$v timeout 600 python bench.py --webvid-videos 0 --train-videos 0 --cpu-sample 8 > $OUT/bench_$v.json 2> $OUT/bench_$v.err; echo "bench '$v' exit $?" | tee -a $OUT/summary.txt
  python -c "
import json; d=json.load(open('$OUT/bench_$v.json')); print('variant [$v]', d['value'], d['ms_per_step'], d['e2e_roofline_frac'], d['e2e']['value'], d['clocks']['sm_mhz']); print(d['roofline']['ms_by_kernel_class'], d['roofline']['achieved']); print([ (r['N'],r['K'],r['tflops']) for r in d['roofline']['top_shapes'][:4]])"
done
cat $OUT/summary.txt
