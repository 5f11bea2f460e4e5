#!/bin/bash
# Scaling session on an 8-GPU box (features sharding, fused hop + NVLink scatter): eager / CUDA graph / NCCL fence.
TAG=${1:-r1s}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # n extra-flags name
  local n=$1 extra=$2 name=$3
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 $extra > $OUT/bench_${n}_$name.log 2>$OUT/bench_${n}_$name.err
  echo "== n=$n $name exit $?"; tail -1 $OUT/bench_${n}_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
    print('   ms/step %.3f  value %.3e  e2e_ms %.2f  hop_ms %s share %s  %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], r.get('ms_per_launch'), r.get('kernel_share_of_step'), d['config']['parallelism']))
except Exception as e: print('   parse failed', e)
" || tail -5 $OUT/bench_${n}_$name.err
}
run 8 "" flags
run 8 "--graph" graph
run 8 "--fence nccl" nccl
run 4 "" flags
