#!/bin/bash
# One multi-GPU box visit: bench.py exactly as the driver launches it (default grid, then alternative grids), then the
# multi-process parity + timing script.  Usage: bash scripts/gpu_round_multi.sh N TAG [ALT_GRID ...]
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29501 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
p=29502
for g in "$@"; do
  timeout 300 $RUN --master-port $p bench.py --gpus $N --steps 30 --warmup 5 --grid $g --no-e2e > gpurun_out/bench_${TAG}_n${N}_$g.json 2>> gpurun_out/bench_${TAG}_n$N.err
  p=$((p+1))
done
MGPU_QUICK=${MGPU_QUICK:-0} timeout 900 $RUN --master-port 29511 tests/multi_gpu/mgpu_check.py > gpurun_out/mgpu${N}_$TAG.log 2>&1
grep -h '"metric"' gpurun_out/bench_${TAG}_n$N*.json | cut -c1-400
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/mgpu${N}_$TAG.log | tail -70
