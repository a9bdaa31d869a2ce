#!/bin/bash
# Round-2 GPU session: parity tests, arithmetic divergence, bench, ncu launch list + full capture.
# Usage (under gpurun, from the repo root): bash scripts/r2_gpu1.sh <tag> [skip_ncu]
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/env.txt 2>&1
nproc >> $OUT/env.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -15 $OUT/pytest.log
fi
if [ -z "$SKIP_ARITH" ]; then
timeout 600 python scripts/arith_divergence.py c1 h0 c3 --out $OUT/arith_divergence.json > $OUT/arith.log 2>&1; echo "arith rc=$?"
tail -3 $OUT/arith.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"
cat $OUT/bench_n1.json; tail -5 $OUT/bench_n1.err
if [ -z "$SKIP_NCU" ]; then
# (graphs off under ncu: the launch indices below count the eager launches of the warm-up / timed steps)
DMGS_BENCH_GRAPHS=0 DMGS_BENCH_VIEWS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 140 --csv \
    --log-file $OUT/launches_h0.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1
DMGS_BENCH_GRAPHS=0 DMGS_BENCH_VIEWS=1 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'blend_bwd|blend_fwd|preprocess_bwd|preprocess_fwd|tile_place|tile_count|radix_scatter' -s 40 -c 14 -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_full_h0_raw.csv 2>/dev/null
# sh_grad_expand with the records of all 8 views of a step
DMGS_BENCH_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sh_grad_expand' -s 3 -c 1 -o $OUT/expand -f \
    python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline > $OUT/ncu_expand.log 2>&1
ncu -i $OUT/expand.ncu-rep --page raw --csv > $OUT/ncu_expand_raw.csv 2>/dev/null
# kernel timeline of the multi-stream step (torch.profiler / CUPTI): 4 streams and 1 stream
python scripts/trace_step.py $OUT 4 1 2 > $OUT/trace_s4.log 2>&1
python scripts/trace_step.py $OUT 1 0 2 > $OUT/trace_s1.log 2>&1
ls -la $OUT
fi
