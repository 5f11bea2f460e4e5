#!/bin/bash
# Multi-GPU session: NCCL parity tests + bench in both shardings.  Usage: bash tools/gpu_multi.sh <tag> <ngpus>
TAG=${1:-r1m}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest distributed"; timeout 600 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/pytest_dist.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_dist.log
echo "== bench 1 gpu"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_1.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_1.log | cut -c1-400
for MODE in nodes features; do
  echo "== bench $NG gpus $MODE"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 --mode $MODE > $OUT/bench_${NG}_$MODE.log 2>&1; echo "exit $?"; tail -2 $OUT/bench_${NG}_$MODE.log | cut -c1-600
done
