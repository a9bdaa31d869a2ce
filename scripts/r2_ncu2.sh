#!/bin/bash
# ncu --set full of the persistent blend kernels (residency 8 and 6) and of sh_grad_expand with 8 views' records.
TAG=${1:-ncu2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for r in 8 6; do
DMGS_BENCH_GRAPHS=0 DMGS_BLEND_FWD_RESIDENCY=$r DMGS_BLEND_BWD_RESIDENCY=$r DMGS_BENCH_VIEWS=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'blend_bwd|blend_fwd' -s 8 -c 2 -o $OUT/blend_r$r -f \
    python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline > $OUT/ncu_blend_r$r.log 2>&1
ncu -i $OUT/blend_r$r.ncu-rep --page raw --csv > $OUT/blend_r${r}_raw.csv 2>/dev/null
done
DMGS_BENCH_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'sh_grad_expand' -s 3 -c 1 -o $OUT/expand -f \
    python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline > $OUT/ncu_expand.log 2>&1
ncu -i $OUT/expand.ncu-rep --page raw --csv > $OUT/expand_raw.csv 2>/dev/null
ncu -i $OUT/expand.ncu-rep --page source --csv > $OUT/expand_source.csv 2>/dev/null
ls -la $OUT
