#!/bin/bash
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for MB in 0 32 64 96; do
  echo "== sweep3 persist=$MB"; SWEEP_PERSIST_MB=$MB SWEEP_R2=3 timeout 300 tools/bin/spmm_sweep 1000000 32 64 8 > $OUT/sweep3_persist$MB.log 2>&1; echo "exit $?"; head -4 $OUT/sweep3_persist$MB.log | tail -2; tail -14 $OUT/sweep3_persist$MB.log
done
echo "== pytest new (fused layer, dmma)"; timeout 900 python -m pytest tests/test_fused_layer.py tests/test_gpu_parity.py tests/test_pooling.py -q -m gpu > $OUT/pytest_new.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_new.log
echo "== bench f64"; timeout 600 python bench.py --dtype f64 --steps 10 --no-cpu-baseline --configs '' > $OUT/bench_er1m_f64.log 2>$OUT/bench_f64.err; echo "exit $?"; tail -1 $OUT/bench_er1m_f64.log | cut -c1-1600
