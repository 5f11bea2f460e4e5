#!/bin/bash
# One GPU session: parity tests, smoke, bench (all single-GPU workloads), ncu launch list + full captures of the hop and
# contraction kernels.   Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "exit $?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py > $OUT/bench.log 2>$OUT/bench.err; echo "exit $?"; tail -1 $OUT/bench.log | cut -c1-2400
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 > $OUT/bench_reference.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_reference.log | cut -c1-300
for WL in cfg2 cfg3 cfg4 sbm1m; do
  echo "== bench $WL"; timeout 600 python bench.py --workload $WL --steps 10 --no-cpu-baseline > $OUT/bench_$WL.log 2>/dev/null; echo "exit $?"; tail -1 $OUT/bench_$WL.log | cut -c1-330
done
echo "== bench f64"; timeout 600 python bench.py --dtype f64 --steps 10 --no-cpu-baseline > $OUT/bench_er1m_f64.log 2>/dev/null; echo "exit $?"; tail -1 $OUT/bench_er1m_f64.log | cut -c1-330
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm_hop|tc_contract|transpose_kernel|pack_taps|tap_|bias_grad" -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "exit $?"
echo "== ncu full hop"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_hop -s 8 -c 1 -o $OUT/prof_hop python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_hop.log 2>&1; echo "exit $?"
echo "== ncu full tc"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_contract -s 2 -c 1 -o $OUT/prof_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_tc.log 2>&1; echo "exit $?"
ls $OUT
