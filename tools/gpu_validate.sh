#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; grep -a "full-size parity\|GRNN\|passed\|failed\|FAILED" $OUT/pytest_gpu.log | tail -14
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "exit $?"; tail -2 $OUT/smoke.log
echo "== sweep narrow C=16"; SWEEP_R2=1 timeout 300 tools/bin/spmm_sweep 1000000 32 16 8 > $OUT/sweep_narrow_c16.log 2>&1; echo "exit $?"; tail -14 $OUT/sweep_narrow_c16.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.log 2>$OUT/bench.err; echo "exit $?"; tail -1 $OUT/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('er1m ms', d['ms_per_step'], 'hop', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'], 'dram_frac', d['roofline'].get('dram_frac'), 'parity', d.get('parity_max_rel'), 'fwd_bwd', d['fwd_bwd']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'ok', d.get('parity_ok'))
for k, v in d.get('configs', {}).items():
    print(k, {a: v.get(a) for a in ('ms_per_step', 'value', 'parity_max_rel', 'hop_ms', 'error')}, (v.get('edge_variant_part') or {}).get('ms'))
print(d.get('cpu_baseline'))
"; tail -3 $OUT/bench.err
echo "== ncu full tc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 2 -c 1 -o $OUT/prof_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check --configs '' > $OUT/ncu_full_tc.log 2>&1; echo "exit $?"
