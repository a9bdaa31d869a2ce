#!/bin/bash
# Persistent blend kernels: parity tests, then `value` over the residency (CTAs per SM) of the two blend kernels.
TAG=${1:-resid}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest.log
fi
for cfg in ${CFGS:-"8,8" "7,7" "6,6" "5,5" "6,7" "7,6"}; do
  f=${cfg%,*}; b=${cfg#*,}
  DMGS_BLEND_FWD_RESIDENCY=$f DMGS_BLEND_BWD_RESIDENCY=$b timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline \
      > $OUT/bench_f${f}_b$b.json 2> $OUT/bench_f${f}_b$b.err
  echo "fwd=$f bwd=$b rc=$? $(python -c "
import json,sys
d=json.loads(open('$OUT/bench_f${f}_b$b.json').read().strip().splitlines()[-1])
st=d['stages']
print(round(d['value'],1),'frames/s',round(d['ms_per_step'],3),'ms/step; single-stream stages', {k: v['ms'] for k,v in st.items()})
" 2>&1 | tail -1)"
  tail -2 $OUT/bench_f${f}_b$b.err
done
