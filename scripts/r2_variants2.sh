#!/bin/bash
# variant .so (scripts/variants.sh) x runtime residency: "name:residency" pairs
TAG=${1:-var2}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for vr in "$@"; do
  v=${vr%:*}; r=${vr#*:}
  lib=build/variants/libdmgs_$v.so
  [ "$v" = base ] && lib=dmgs_b200/libdmgs_raster.so
  DMGS_RASTER_LIB=$lib DMGS_VIEW_BLEND_RESIDENCY=$r DMGS_BLEND_FWD_RESIDENCY=$r DMGS_BLEND_BWD_RESIDENCY=$r timeout 300 python bench.py --steps 8 --warmup 3 --quick --no-cpu-baseline > $OUT/${v}_$r.json 2> $OUT/${v}_$r.err
  python - "$v:$r" $OUT/${v}_$r.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 1), {k: v["ms"] for k, v in d["stages"].items() if "blend" in k})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
