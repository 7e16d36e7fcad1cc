#!/usr/bin/env bash
# One GPU-box visit (round 2): stages selected by STAGES="a b c"; outputs under gpurun_out/$VISIT/.
#   lockstep  scripts/debug_lockstep.py          pool   scripts/sweep_k1_order.py
#   tests     pytest -m gpu                      bench  bench.py (N=1, all configs)
#   parity    scripts/parity_report.py           ncu_k1 / ncu_k2 / ncu_k3  one --set full capture each
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${VISIT:-r2}
mkdir -p $O
for st in ${STAGES:-tests bench}; do
  t0=$(date +%s)
  case $st in
    lockstep) timeout 300 python scripts/debug_lockstep.py > $O/lockstep.log 2>&1 ;;
    pool) timeout 400 python scripts/sweep_k1_order.py ${POOL_ARGS:-} > $O/sweep_k1_order.jsonl 2> $O/sweep_k1_order.err ;;
    tests) timeout ${TEST_TIMEOUT:-900} python -m pytest tests -q -m gpu -x --durations=10 ${PYTEST_ARGS:-} > $O/pytest_gpu.log 2>&1 ;;
    bench) timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > $O/bench_n1.json 2> $O/bench_n1.err ;;
    parity) timeout 1200 python scripts/parity_report.py --instances ${PARITY_B:-256} --out $O/parity_r2.json > $O/parity.log 2>&1 ;;
    configs) timeout 600 python scripts/bench_configs.py ${CONFIG_BUDGET:-30} ${RUN_CONFIGS:-3,4a} > $O/bench_configs.log 2>&1 ;;
    ncu_k1) timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_loop_kernel --launch-skip 1 -c 1 \
              -o $O/prof_k1 -f python scripts/sweep_k1_order.py --sizes ${NCU_K1_SIZE:-1048576} --steps 1 > $O/ncu_k1.log 2>&1
            [ -f $O/prof_k1.ncu-rep ] && python scripts/summarise_ncu.py $O/prof_k1.ncu-rep $O/k1.ncu.txt > /dev/null 2>> $O/ncu_k1.log ;;
    ncu_k2) timeout 500 ncu --set full --clock-control none --import-source on -k regex:k2_loop_kernel --launch-skip 1 -c 1 \
              -o $O/prof_k2 -f python scripts/bench_configs.py 1000 3 ${NCU_K2_B:-4096} > $O/ncu_k2.log 2>&1
            [ -f $O/prof_k2.ncu-rep ] && python scripts/summarise_ncu.py $O/prof_k2.ncu-rep $O/k2.ncu.txt > /dev/null 2>> $O/ncu_k2.log ;;
    ncu_k3) timeout 500 ncu --set full --clock-control none --import-source on -k regex:k3_loop_kernel --launch-skip 1 -c 1 \
              -o $O/prof_k3 -f python scripts/bench_configs.py 1000 4a ${NCU_K3_B:-2048} > $O/ncu_k3.log 2>&1
            [ -f $O/prof_k3.ncu-rep ] && python scripts/summarise_ncu.py $O/prof_k3.ncu-rep $O/k3.ncu.txt > /dev/null 2>> $O/ncu_k3.log ;;
    launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
              python bench.py --steps 2 --warmup 1 --no-cpu-baseline ${BENCH_ARGS:-} > $O/bench_under_ncu.log 2>&1 ;;
    *) echo "unknown stage $st" ;;
  esac
  echo "$st rc=$? $(( $(date +%s) - t0 )) s" >> $O/stages.txt
done
cat $O/stages.txt
[ -f $O/lockstep.log ] && tail -n 8 $O/lockstep.log
[ -f $O/sweep_k1_order.jsonl ] && cut -c1-260 $O/sweep_k1_order.jsonl
[ -f $O/pytest_gpu.log ] && tail -n 15 $O/pytest_gpu.log
[ -f $O/bench_n1.json ] && cut -c1-1500 $O/bench_n1.json && tail -n 5 $O/bench_n1.err
exit 0
