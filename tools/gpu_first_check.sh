#!/bin/bash
# First GPU call of a round: run the GPU legs that were written without GPU access (static-GSO recurrent layers,
# batch-/time-varying filter, and — with >= 2 GPUs — the multi-GPU backward), each file on its own so that one failure
# does not hide the others, then the full GPU suite and the smoke test.
#   Usage (repo root, under gpurun):  bash tools/gpu_first_check.sh [tag]
TAG=${1:-first}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
export B200GF_STRICT_WIDEN=1   # report the first hardware run of these legs as ordinary pass / fail
for T in test_widen_recurrent test_widen_delayed test_widen_sparse_inputs; do
  echo "== $T"; timeout 600 python -m pytest tests/$T.py -q -m gpu > $OUT/$T.log 2>&1; echo "exit $?"; tail -15 $OUT/$T.log
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== test_widen_distributed (NCCL, forward + backward)"; timeout 900 python -m pytest tests/test_widen_distributed.py -q -m gpu > $OUT/test_distributed.log 2>&1; echo "exit $?"; tail -15 $OUT/test_distributed.log
fi
echo "== full GPU suite"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "exit $?"; tail -2 $OUT/smoke.log
# next-kernel probe (standalone, not part of the library): tensor-core tap gradient, correctness + timing at R = 1M
echo "== tapgrad_tc_probe"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tapgrad_tc_probe tools/tapgrad_tc_probe.cu > $OUT/probe_build.log 2>&1 \
  && timeout 120 /tmp/tapgrad_tc_probe time > $OUT/tapgrad_tc_probe.log 2>&1; echo "exit $?"; tail -8 $OUT/tapgrad_tc_probe.log
echo "== contract_f64_probe"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/contract_f64_probe tools/contract_f64_probe.cu >> $OUT/probe_build.log 2>&1 \
  && timeout 120 /tmp/contract_f64_probe time > $OUT/contract_f64_probe.log 2>&1; echo "exit $?"; tail -8 $OUT/contract_f64_probe.log
