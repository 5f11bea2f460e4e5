#!/bin/bash
# ncu --set full of the edge-variant step kernel (k = 1 of cfg4ev) and of the tap-gradient kernel (er1m backward)
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== ncu ev step"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $OUT/prof_ev_step python bench.py --workload tiny --steps 3 --configs cfg4ev --no-cpu-baseline --no-check > $OUT/ncu_ev.log 2>&1; echo "exit $?"
echo "== ncu tap_grad"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tap_grad_multi -s 1 -c 1 -o $OUT/prof_tap_grad python bench.py --steps 3 --configs '' --no-cpu-baseline --no-check > $OUT/ncu_tg.log 2>&1; echo "exit $?"
ls -la $OUT
