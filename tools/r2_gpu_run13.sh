#!/bin/bash
OUT=gpurun_out/r2_run13
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_kernels.py tests/test_gpu_encoder.py -m gpu -q -x --tb=short > $OUT/pytest_part.log 2>&1; echo "pytest part exit $?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_part.log
timeout 300 python tools/membound_bench.py 2> $OUT/membound.err | grep "layernorm\|ln_bwd\|adamw" > $OUT/membound.jsonl; cat $OUT/membound.jsonl
FITCLIP_VARIANT=lnb0 timeout 300 python tools/membound_bench.py 2> $OUT/membound0.err | grep "ln_bwd" > $OUT/membound_lnb0.jsonl; cat $OUT/membound_lnb0.jsonl
for v in "" lnb0; do
FITCLIP_VARIANT=$v timeout 300 python tools/train_step.py --videos 512 --steps 3 --warmup 2 > $OUT/train_step_$v.json 2> $OUT/train_step_$v.err
python -c "
import json; d=json.load(open('$OUT/train_step_$v.json')); print('[$v]', d['ms_per_step'], {k:(round(x['ms'],1), round(x['gbs'])) for k,x in d['profiled_ms_by_kernel_class'].items() if k in ('layernorm','layernorm_bwd','adamw','quickgelu','quickgelu_bwd','colsum')})"
done
cat $OUT/summary.txt
