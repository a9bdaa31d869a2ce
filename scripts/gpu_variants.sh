#!/bin/bash
# Times kernel variants (scripts/variants.sh) with the bench's stage timing: one line per variant.
TAG=${1:-variants}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  lib=build/variants/libdmgs_$v.so
  [ "$v" = base ] && lib=dmgs_b200/libdmgs_raster.so
  DMGS_RASTER_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --quick > $OUT/$v.json 2> $OUT/$v.err
  python - "$v" $OUT/$v.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "value", round(d["value"], 1), {k: v["ms"] for k, v in d["stages"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
