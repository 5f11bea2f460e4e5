#!/bin/bash
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== sweep2 C=64"; SWEEP_R2=2 timeout 600 tools/bin/spmm_sweep 1000000 32 64 10 > $OUT/sweep2_c64.log 2>&1; echo "exit $?"; tail -22 $OUT/sweep2_c64.log
