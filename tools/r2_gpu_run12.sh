#!/bin/bash
OUT=gpurun_out/r2_run12
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_kernels.py tests/test_gpu_encoder.py tests/test_gpu_training.py tests/test_reference_golden.py tests/test_reference_training_golden.py -m gpu -q -x --tb=short > $OUT/pytest_part.log 2>&1; echo "pytest part exit $?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_part.log
timeout 300 python tools/membound_bench.py > $OUT/membound.jsonl 2> $OUT/membound.err; echo "membound exit $?" | tee -a $OUT/summary.txt
cat $OUT/membound.jsonl
timeout 900 python bench.py --webvid-videos 0 > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
python -c "
import json; d=json.load(open('$OUT/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e_roofline_frac'], d['e2e']['value'], d['clocks']['sm_mhz']); print(d['roofline']['ms_by_kernel_class'], d['roofline']['achieved']); print(d['extra']['train_step'])"
cat $OUT/summary.txt
