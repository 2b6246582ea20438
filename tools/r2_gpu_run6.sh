#!/bin/bash
OUT=gpurun_out/r2_run6
mkdir -p $OUT
timeout 120 python tools/attention_bench.py >> $OUT/attention_bench.log 2>&1; echo "attention_bench exit $?" | tee -a $OUT/summary.txt
cat $OUT/attention_bench.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q -x --tb=short > $OUT/pytest_part.log 2>&1; echo "pytest part exit $?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_part.log
timeout 600 python tools/pass_size_sweep.py 384 480 500 512 640 768 960 1280 1920 3840 > $OUT/pass_size_sweep.txt 2>&1; echo "sweep exit $?" | tee -a $OUT/summary.txt
cat $OUT/pass_size_sweep.txt
timeout 600 python tools/geometry_bench.py clip_vit_l_14 clip_vit_l_14_336px > $OUT/geometry_bench.jsonl 2> $OUT/geometry_bench.err; echo "geometry exit $?" | tee -a $OUT/summary.txt
cat $OUT/geometry_bench.jsonl
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 3 --webvid-videos 0 --train-videos 0 --cpu-sample 8 > $OUT/ncu_bench.log 2>&1; echo "ncu launches exit $?" | tee -a $OUT/summary.txt
cat $OUT/summary.txt
