#!/bin/bash
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for C in 8 16; do
  echo "== sweep narrow C=$C"; SWEEP_R2=1 timeout 300 tools/bin/spmm_sweep 1000000 32 $C 8 > $OUT/sweep_narrow_c$C.log 2>&1; echo "exit $?"; tail -14 $OUT/sweep_narrow_c$C.log
done
echo "== pytest ev + recurrent (graphed GRNN, chain on)"; timeout 900 python -m pytest tests/test_evgf.py tests/test_widen_recurrent.py tests/test_cabi.py -q -m gpu -s > $OUT/pytest_a.log 2>&1; echo "exit $?"; grep -a "GRNN\|passed\|failed" $OUT/pytest_a.log | tail -6
echo "== graphed GRNN, chain off"; B200GF_CHAIN_MAX_BYTES=0 timeout 600 python -m pytest tests/test_widen_recurrent.py -q -m gpu -s -k graphed > $OUT/pytest_nochain.log 2>&1; echo "exit $?"; grep -a "GRNN\|passed\|failed" $OUT/pytest_nochain.log | tail -4
echo "== tc raw-hi experiment"; B200GF_TC_RAWHI=1 timeout 600 python -m pytest tests/test_cabi.py -q -m gpu -k tensor_core > $OUT/pytest_rawhi.log 2>&1; echo "exit $?"; tail -4 $OUT/pytest_rawhi.log
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench.log 2>$OUT/bench.err; echo "exit $?"; tail -1 $OUT/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('er1m ms', d['ms_per_step'], 'hop', d['roofline']['ms_per_launch'], 'parity', d.get('parity_max_rel'), 'fwd_bwd', d['fwd_bwd']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k, v in d.get('configs', {}).items():
    print(k, {a: v.get(a) for a in ('ms_per_step', 'value', 'parity_max_rel', 'hop_ms', 'error')}, (v.get('edge_variant_part') or {}).get('ms'), ((v.get('edge_variant_part') or {}).get('roofline') or {}).get('frac'))
"; tail -3 $OUT/bench.err
echo "== bench rawhi (contraction only matters)"; B200GF_TC_RAWHI=1 timeout 600 python bench.py --steps 10 --configs '' --no-cpu-baseline > $OUT/bench_rawhi.log 2>/dev/null; echo "exit $?"; tail -1 $OUT/bench_rawhi.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('rawhi: ms', d['ms_per_step'], 'parity', d.get('parity_max_rel'))"
