#!/bin/bash
# What the driver runs at round end, in the same order: GPU test suite, smoke(), the reference arm, the bench.
OUT=gpurun_out/r2_final
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?" | tee -a $OUT/summary.txt
cut -c1-400 $OUT/bench_reference.json
timeout 900 python bench.py --save-webvid-record $OUT/webvid_1gpu_record.json > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
python -c "
import json; d=json.load(open('$OUT/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e_roofline_frac'], d['e2e']['value'], d['extra']['e2e_uint8']['value'], d['clocks']); print(d['roofline']['ms_by_kernel_class'], d['roofline']['achieved'], d['roofline']['frac']); w=d['extra']['webvid']; print({k:w[k] for k in ('seconds','videos_per_s','roofline_frac','sampled_rows_match_oracle')}); print(d['extra']['train_step']['ms_per_step'], d['extra']['train_step']['achieved_tflops']); print(d['cpu_baseline'])"
timeout 300 python tools/membound_bench.py > $OUT/membound.jsonl 2> $OUT/membound.err; cat $OUT/membound.jsonl | cut -c1-160
cat $OUT/summary.txt
