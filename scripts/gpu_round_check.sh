#!/usr/bin/env bash
# One GPU-box visit: A/B sweeps of the specialised builds (bitwise check against the general kernels), the GPU suite
# (FULL=1 includes the one 2-minute full-size Burgers test), the bench line, the ncu captures of the headline kernel
# and, with RUN_CONFIGS=3,4a,5c, the other BASELINE configs. Outputs under gpurun_out/$VISIT/ (default v2).
# Knobs: VISIT, K1_SPECS, USE_BEST (run bench/ncu with the sweep's fastest build), FULL, RUN_CONFIGS, SKIP_TESTS.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${VISIT:-v2}
mkdir -p $O
timeout 240 python scripts/sweep_k1_spec.py --specs ${K1_SPECS:-0,1,2,3} --out $O/best_spec.txt > $O/sweep_k1_spec.jsonl 2> $O/sweep_k1.err
echo "sweep_k1 rc=$?" >> $O/stages.txt
if [ -n "${USE_BEST:-}" ] && [ -s $O/best_spec.txt ]; then export PDEQ_K1_SPEC=$(cat $O/best_spec.txt); fi
echo "PDEQ_K1_SPEC=${PDEQ_K1_SPEC:-library default}" >> $O/stages.txt
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 240 python scripts/sweep_k2_spec.py 8192 > $O/sweep_k2_spec.jsonl 2> $O/sweep_k2.err
echo "sweep_k2 rc=$?" >> $O/stages.txt
DESELECT="--deselect tests/test_gpu_group_and_smoother.py::test_burgers_d1024_ts1_full_size_dimension"
if [ -n "${FULL:-}" ]; then DESELECT=""; fi
timeout 600 python -m pytest tests -q -m gpu --durations=8 $DESELECT > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/stages.txt
fi
timeout 240 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/stages.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_loop_kernel --launch-skip 1 -c 1 \
  -o $O/prof_k1 -f python scripts/sweep_k1_spec.py --specs ${PDEQ_K1_SPEC:-2} --steps 2 --warmup 1 > $O/ncu_full.log 2>&1
echo "ncu_full rc=$?" >> $O/stages.txt
[ -f $O/prof_k1.ncu-rep ] && python scripts/summarise_ncu.py $O/prof_k1.ncu-rep $O/k1.ncu.txt > /dev/null 2>> $O/ncu_full.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "ncu_launches rc=$?" >> $O/stages.txt
if [ -n "${RUN_CONFIGS:-}" ]; then
  timeout 240 python scripts/bench_configs.py 30 ${RUN_CONFIGS} > $O/bench_configs.log 2>&1
  echo "bench_configs rc=$?" >> $O/stages.txt
fi
cat $O/stages.txt
[ -f $O/pytest_gpu.log ] && tail -n 12 $O/pytest_gpu.log
cut -c1-220 $O/sweep_k1_spec.jsonl; [ -f $O/sweep_k2_spec.jsonl ] && cut -c1-220 $O/sweep_k2_spec.jsonl; cut -c1-400 $O/bench_n1.json
