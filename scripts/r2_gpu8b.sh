#!/bin/bash
# 8-GPU session, trimmed: H0 weak-scaled and C4 (3 M Gaussians, 64 views, 1245x825) at 8 ranks.
TAG=${1:-r2n8b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { # name, N, extra env...
  local name=$1; local N=$2; shift; shift
  timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  echo "$name rc=$?"; python - $OUT/$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), "frames/s", round(d["ms_per_step"], 3), "ms/step; e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["training_step"], 1),
          "; tail", json.dumps({k: v for k, v in (d.get("optimizer_tail") or {}).items() if k != "note"}), "; nrank", d.get("nrank_vs_1rank_rel_err"), d.get("allreduce_check"))
except Exception as e:
    print("failed", e)
PY
  tail -2 $OUT/$name.err | cut -c1-300
}
run bench_h0_n8 8 DMGS_BENCH_WORKLOAD=h0
run bench_c4_n8 8 DMGS_BENCH_WORKLOAD=c4
