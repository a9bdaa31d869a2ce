#!/bin/bash
# Full ncu capture of selected kernels: bash scripts/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-ncu}; RE=${2:-blend}; SKIP=${3:-20}; COUNT=${4:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
DMGS_BENCH_VIEWS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $COUNT \
    -o $OUT/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo rc=$?; tail -3 $OUT/ncu_full.log | cut -c1-300
