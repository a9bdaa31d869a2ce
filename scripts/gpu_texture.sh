#!/bin/bash
# Texture row: timing + full ncu capture of its kernels.  bash scripts/gpu_texture.sh <tag>
TAG=${1:-tex}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/bench_next_rows.py texture > $OUT/texture.jsonl 2> $OUT/texture.err; echo rc=$?
cat $OUT/texture.jsonl
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'texture|tex_' -s 9 -c 8 -o $OUT/prof -f \
    python scripts/bench_next_rows.py texture > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py $OUT/prof.ncu-rep > $OUT/ncu_summary.txt 2>&1
tail -40 $OUT/ncu_summary.txt
fi
