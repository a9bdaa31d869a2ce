#!/bin/bash
# ncu launch list of ~1 frame (cold-cache, serialised): per-kernel device time.
TAG=${1:-launches}
OUT=gpurun_out/$TAG
mkdir -p $OUT
DMGS_BENCH_VIEWS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-150} -c ${COUNT:-60} --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
echo rc=$?
