#!/bin/bash
# round 2, GPU call 1: hop-kernel sweep (v2 / ldg256 / cp.async variants), the two never-run probes, full GPU suite, smoke
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== sweep C=64"; SWEEP_R2=1 timeout 600 tools/bin/spmm_sweep 1000000 32 64 10 > $OUT/sweep_c64.log 2>&1; echo "exit $?"; tail -28 $OUT/sweep_c64.log
echo "== tapgrad_tc_probe"; timeout 120 tools/bin/tapgrad_tc_probe time > $OUT/tapgrad_tc_probe.log 2>&1; echo "exit $?"; tail -12 $OUT/tapgrad_tc_probe.log
echo "== contract_f64_probe"; timeout 120 tools/bin/contract_f64_probe time > $OUT/contract_f64_probe.log 2>&1; echo "exit $?"; tail -12 $OUT/contract_f64_probe.log
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; tail -15 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "exit $?"; tail -3 $OUT/smoke.log
