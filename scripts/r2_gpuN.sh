#!/bin/bash
# Multi-GPU session: bash scripts/r2_gpuN.sh <N> <tag>   (under gpurun --gpus N)
N=${1:-2}
TAG=${2:-r2n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/env.txt 2>&1
nvidia-smi topo -m >> $OUT/env.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer_allreduce.py -x -q > $OUT/pytest_peer.log 2>&1; echo "pytest peer rc=$?"; tail -3 $OUT/pytest_peer.log
run() { # name, extra env/args...
  local name=$1; shift
  timeout 900 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT/$name.json 2> $OUT/$name.err
  echo "$name rc=$?"; cat $OUT/$name.json; tail -3 $OUT/$name.err
}
run bench_h0_n$N DMGS_BENCH_WORKLOAD=h0
run bench_c4_n$N DMGS_BENCH_WORKLOAD=c4
if [ -n "$WITH_NCCL" ]; then run bench_h0_n${N}_nccl DMGS_BENCH_WORKLOAD=h0 DMGS_BENCH_ALLREDUCE=nccl; fi
