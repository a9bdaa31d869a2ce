#!/bin/bash
# Quick GPU iteration: parity tests + one bench line (no CPU baseline) + optional launch list (LAUNCHES=1).
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ -n "$LAUNCHES" ]; then bash scripts/gpu_launches.sh $TAG; fi
