#!/bin/bash
# Scaling session on an 8-GPU box: bench at 1, 2, 4, 8 ranks (features sharding), plus the node sharding at 8.
TAG=${1:-r1s}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # n mode
  local n=$1 mode=$2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${n}_$mode.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 --mode $mode > $OUT/bench_${n}_$mode.log 2>&1
  fi
  echo "== n=$n mode=$mode exit $?"; tail -1 $OUT/bench_${n}_$mode.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   ms/step %.3f  value %.3e  e2e_ms %.2f  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['clocks']))
except Exception as e: print('   parse failed', e)
"
}
run 1 features
run 2 features
run 4 features
run 8 features
run 8 nodes
run 4 nodes
