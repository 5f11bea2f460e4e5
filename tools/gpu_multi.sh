#!/bin/bash
# Multi-GPU session: NCCL parity tests + bench (features sharding fused / NCCL all-to-all, node sharding).
# Usage: bash tools/gpu_multi.sh <tag> <ngpus>
TAG=${1:-r1m}; NG=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest distributed"; timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/pytest_dist.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_dist.log
for CFG in "features" "features --graph" "features --fence nccl" "nodes"; do
  NAME=$(echo $CFG | tr -d ' -')
  echo "== bench $NG gpus $CFG"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 --mode $CFG > $OUT/bench_${NG}_$NAME.log 2>&1; echo "exit $?"
  tail -1 $OUT/bench_${NG}_$NAME.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   ms/step %.3f  value %.3e  e2e_ms %.2f  %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['parallelism']))
except Exception as e: print('   parse failed', e)
"
done
