#!/bin/bash
# Scaling session on an 8-GPU box: bench at 1, 2, 4, 8 ranks (features sharding, fused hop + NVLink scatter),
# plus the NCCL all-to-all variant and the node sharding at 8 for comparison.
TAG=${1:-r1s}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # n mode extra-flags name
  local n=$1 mode=$2 extra=$3 name=$4
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${n}_$name.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 --mode $mode $extra > $OUT/bench_${n}_$name.log 2>&1
  fi
  echo "== n=$n $name exit $?"; tail -1 $OUT/bench_${n}_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('   ms/step %.3f  value %.3e  e2e_ms %.2f  %s  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['parallelism'], d['clocks']['reasons']))
except Exception as e: print('   parse failed', e); 
" || tail -5 $OUT/bench_${n}_$name.log
}
run 1 features "" fused
run 2 features "" fused
run 4 features "" fused
run 8 features "" fused
run 8 features "--no-fused" nccl_a2a
