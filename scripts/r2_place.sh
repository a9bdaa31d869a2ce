#!/bin/bash
# placement smem budget x blend residency: `value` and the single-stream binning time
TAG=${1:-place}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "$@"; do
  kb=${cfg%:*}; r=${cfg#*:}
  DMGS_PLACE_SMEM_KB=$kb DMGS_VIEW_BLEND_RESIDENCY=$r timeout 300 python bench.py --steps 8 --warmup 3 --quick --no-cpu-baseline > $OUT/p${kb}_r$r.json 2> $OUT/p${kb}_r$r.err
  python - "$kb:$r" $OUT/p${kb}_r$r.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 1), round(d["ms_per_step"],3), {k: v["ms"] for k, v in d["stages"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
  tail -1 $OUT/p${kb}_r$r.err | cut -c1-200
done
