#!/bin/bash
# round 2 multi-GPU session.  Usage (repo root, under gpurun --gpus N):  bash tools/r2_multi.sh <tag> <ngpus> [quick]
TAG=${1:-r2m}
NG=${2:-2}
QUICK=${3:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
if [ "$NG" = "2" ] && [ -z "$QUICK" ]; then
  echo "== pytest NCCL world 2 (forward + backward, fused node / feature paths)"
  timeout 900 python -m pytest tests/test_distributed.py tests/test_widen_distributed.py -q -m gpu -x > $OUT/pytest_dist.log 2>&1; echo "exit $?"; tail -6 $OUT/pytest_dist.log
fi
run() {  # name, args...
  local name=$1; shift
  echo "== bench $name: $*"
  timeout 600 $TR bench.py --gpus $NG "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "exit $?"
  tail -1 $OUT/bench_$name.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    st = d.get('selftest', {})
    print('  ms/step %.3f  e2e %.3f  value %.3e  parity %s  mode: %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d.get('parity_max_rel'), d['config']['parallelism']))
    print('  probed', d.get('modes_probed_ms'), ' fwd_bwd', (d.get('fwd_bwd') or {}).get('ms_per_step'), ' hop ms', (d.get('roofline') or {}).get('ms_per_launch'), ' launches', d.get('gpu_launches'))
    for k, v in st.items(): print('  selftest', k, {a: ('%.1e' % b if isinstance(b, float) else b) for a, b in v.items()})
    if d.get('graph_errors'): print('  graph_errors', d['graph_errors'])
except Exception as e:
    print('  (no JSON line)', e)
"
  tail -3 $OUT/bench_$name.err | cut -c1-300
}
run auto
if [ -z "$QUICK" ]; then
  run nodes_ipc --mode nodes --symm ipc --no-selftest
  run features --mode features --no-selftest
  run nodes_nograph --mode nodes --no-graph --no-selftest --no-bwd
  run er2m --workload er2m --mode nodes --no-selftest
else
  run nodes --mode nodes --no-selftest
  run er2m --workload er2m --mode auto --no-selftest
fi
ls $OUT
