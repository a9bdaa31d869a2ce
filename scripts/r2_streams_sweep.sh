#!/bin/bash
# `value` (bench.py --quick) over the number of view streams of multiview.ViewStreams.
for s in 3 4 5 6 8; do DMGS_BENCH_STREAMS=$s timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams=$s', round(d['value'],1), round(d['ms_per_step'],3))"; done
