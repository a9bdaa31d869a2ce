#!/bin/bash
# CUDA-graph views: the new GPU test, then `value` with and without graphs.
TAG=${1:-graphs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_renderer.py -m gpu -x -q -k "captured or view_streams or async_binning" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 $OUT/pytest.log
for cfg in "4 0" "4 1" "8 1" "2 1"; do
  set -- $cfg
  DMGS_BENCH_STREAMS=$1 DMGS_BENCH_GRAPHS=$2 timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline \
      > $OUT/bench_s$1_g$2.json 2> $OUT/bench_s$1_g$2.err
  echo "streams=$1 graphs=$2 rc=$? $(python -c "
import json,sys
d=json.loads(open('$OUT/bench_s$1_g$2.json').read().strip().splitlines()[-1])
print(round(d['value'],1),'frames/s',round(d['ms_per_step'],3),'ms/step host',d.get('host_enqueue_ms_per_step'),d['config'].get('cuda_graphs'),d['config'].get('steps_repeated_after_overflow'))
" 2>&1 | tail -1)"
  tail -3 $OUT/bench_s$1_g$2.err
done
