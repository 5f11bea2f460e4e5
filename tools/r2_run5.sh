#!/bin/bash
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (ev, cabi, parity, fused layer, widen)"; timeout 1500 python -m pytest tests/test_evgf.py tests/test_cabi.py tests/test_gpu_parity.py tests/test_fused_layer.py tests/test_widen_recurrent.py tests/test_widen_delayed.py -q -m gpu -x > $OUT/pytest.log 2>&1; echo "exit $?"; tail -8 $OUT/pytest.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.log 2>$OUT/bench.err; echo "exit $?"; tail -1 $OUT/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('er1m ms', d['ms_per_step'], 'parity', d.get('parity_max_rel'), 'fwd_bwd', d['fwd_bwd']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k, v in d.get('configs', {}).items():
    print(k, {a: v.get(a) for a in ('ms_per_step', 'value', 'parity_max_rel', 'hop_ms', 'error')}, v.get('edge_variant_part'))
"; tail -3 $OUT/bench.err
for MB in 0 64; do
  echo "== sweep3 persist=$MB"; SWEEP_PERSIST_MB=$MB SWEEP_R2=3 timeout 300 tools/bin/spmm_sweep 1000000 32 64 8 > $OUT/sweep3_persist$MB.log 2>&1; echo "exit $?"; sed -n 2,4p $OUT/sweep3_persist$MB.log; tail -13 $OUT/sweep3_persist$MB.log
done
echo "== sweep4 column passes"; SWEEP_R2=4 timeout 300 tools/bin/spmm_sweep 1000000 32 64 8 > $OUT/sweep4_passes.log 2>&1; echo "exit $?"; tail -10 $OUT/sweep4_passes.log
