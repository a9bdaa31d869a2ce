#!/bin/bash
# Round-2 supplementary measurements: CUB library-stage comparator, blend culling statistics, 4K radix-fallback timing.
TAG=${1:-extras}
OUT=gpurun_out/$TAG
mkdir -p $OUT
scripts/_cub_stage 1000000 7376946 2500 20 > $OUT/cub_stage.jsonl 2>&1
scripts/_cub_stage 1000000 2305801 4056 20 >> $OUT/cub_stage.jsonl 2>&1
scripts/_cub_stage 100000 724386 2500 20 >> $OUT/cub_stage.jsonl 2>&1
cat $OUT/cub_stage.jsonl
timeout 300 python scripts/blend_stats.py h0 c1 c3 --out $OUT/blend_stats.json > $OUT/blend_stats.log 2>&1; echo "stats rc=$?"
timeout 600 python scripts/bench_configs.py 4k > $OUT/configs_4k.jsonl 2> $OUT/configs_4k.err; echo "4k rc=$?"
cat $OUT/configs_4k.jsonl; tail -3 $OUT/configs_4k.err
