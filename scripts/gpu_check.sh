#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list and full captures of the top kernels.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/env.txt 2>&1
nproc >> $OUT/env.txt; ls baseline/_ref >> $OUT/env.txt 2>&1
python -c "import diff_gaussian_rasterization, sys; print(diff_gaussian_rasterization.__file__)" >> $OUT/env.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
if [ -z "$SKIP_NCU" ]; then
DMGS_BENCH_VIEWS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
DMGS_BENCH_VIEWS=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'blend_bwd|blend_fwd|preprocess_bwd|preprocess_fwd|radix_scatter|emit_inst' -s 40 -c 12 -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
