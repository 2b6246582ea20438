#!/bin/bash
# usage: bash tools/r2_gpu_multi.sh N   (on a box with N GPUs): NCCL rank-equality tests + bench at N ranks
N=${1:-2}
OUT=gpurun_out/r2_multi_${N}gpu
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --tb=short > $OUT/pytest_multi.log 2>&1; echo "pytest multi exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("$OUT/bench_${N}gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus","ranks_equal_single_gpu") if k in d})
print(d.get("ranks_equal_detail"))
w=d["extra"]["webvid"]; print({k:w[k] for k in w if k in ("seconds","videos_per_s","roofline_frac","strong_scaling_efficiency","metrics_equal_single_gpu","sampled_rows_match_oracle","metrics")})
PY
cat $OUT/summary.txt
