#!/bin/bash
# `value` over the blend residency ViewStreams applies (final build: 128 KB placement budget, 128-thread per-Gaussian CTAs).
for r in 7 6 5; do DMGS_VIEW_BLEND_RESIDENCY=$r timeout 300 python bench.py --steps 16 --warmup 3 --quick --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('residency=$r', round(d['value'],1), round(d['ms_per_step'],3))"; done
