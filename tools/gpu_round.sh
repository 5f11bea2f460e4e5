#!/bin/bash
# One GPU session: parity tests, smoke, kernel sweep, bench, ncu launch list + full capture of the hop kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; tail -15 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "exit $?"; tail -4 $OUT/smoke.log
echo "== sweep"; timeout 300 tools/spmm_sweep 1000000 32 64 10 > $OUT/sweep_c64.log 2>&1; echo "exit $?"; cat $OUT/sweep_c64.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2>&1; echo "exit $?"; tail -3 $OUT/bench.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "exit $?"
echo "== ncu full hop"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_hop -s 8 -c 2 -o $OUT/prof_hop python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "exit $?"
ls -la $OUT
