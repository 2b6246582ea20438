#!/bin/bash
# In-step A/B of attention variants (the step runs power-capped at ~1.29 GHz, where the attention kernel is SM-bound, not
# HBM-bound as it is in isolation at 1.9 GHz): bench's cfg2 step with per-class kernel times.
#   usage: bash tools/attention_poly_ab.sh "VARIANT_OR_ENV ..."     entries: a `make VARIANT=` name, or KEY=VALUE for the environment
mkdir -p gpurun_out/r2_ab
LIST=${1:-"- FC_ATTENTION=tc1 -"}
for v in $LIST; do
  if [[ "$v" == *=* ]]; then export_env="$v"; variant=""; else export_env="FC_NOOP=1"; variant="$v"; fi
  [[ "$variant" == "-" ]] && variant=""
  env $export_env FITCLIP_VARIANT=$variant timeout 300 python bench.py --steps 10 --warmup 3 --webvid-videos 0 --train-videos 0 --cpu-sample 8 > gpurun_out/r2_ab/bench.json 2> gpurun_out/r2_ab/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_ab/bench.json"))
c=d["roofline"]["ms_by_kernel_class"]
print("variant [$v]", round(d["ms_per_step"],2), {k: round(x/10,2) for k,x in c.items()}, d["clocks"]["sm_mhz"])
PY
done
