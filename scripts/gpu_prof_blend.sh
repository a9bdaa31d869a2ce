#!/bin/bash
# ncu --set full capture of the blend kernels (one H0 view) + raw csv.  bash scripts/gpu_prof_blend.sh <tag> [kernel regex]
TAG=${1:-prof}; RE=${2:-'blend_bwd|blend_fwd'}
OUT=gpurun_out/$TAG
mkdir -p $OUT
DMGS_BENCH_VIEWS=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"$RE" -s ${SKIP:-8} -c ${COUNT:-2} -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 3 --quick > $OUT/ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
ls -la $OUT
