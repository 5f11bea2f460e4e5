#!/bin/bash
# round 2, single-GPU session: GPU suite (incl. full-size oracle parity), bench default (er1m + cfg2/3/4 + parity), f64,
# ncu launch list + --set full of the hop kernel.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -s > $OUT/pytest_gpu.log 2>&1; echo "exit $?"; grep -a "full-size parity\|passed\|failed\|Error" $OUT/pytest_gpu.log | tail -12
echo "== bench"; timeout 900 python bench.py > $OUT/bench.log 2>$OUT/bench.err; echo "exit $?"; tail -1 $OUT/bench.log | cut -c1-3000; tail -5 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > $OUT/bench_reference.log 2>&1; echo "exit $?"; tail -1 $OUT/bench_reference.log | cut -c1-900
echo "== bench f64"; timeout 600 python bench.py --dtype f64 --steps 10 --no-cpu-baseline --configs '' > $OUT/bench_er1m_f64.log 2>$OUT/bench_f64.err; echo "exit $?"; tail -1 $OUT/bench_er1m_f64.log | cut -c1-600
echo "== bench sbm1m"; timeout 600 python bench.py --workload sbm1m --steps 10 --no-cpu-baseline --configs '' > $OUT/bench_sbm1m.log 2>/dev/null; echo "exit $?"; tail -1 $OUT/bench_sbm1m.log | cut -c1-400
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check --configs '' > $OUT/ncu_launches.log 2>&1; echo "exit $?"
echo "== ncu full hop"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_hop -s 8 -c 1 -o $OUT/prof_hop python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-check --configs '' > $OUT/ncu_full_hop.log 2>&1; echo "exit $?"
ls $OUT
