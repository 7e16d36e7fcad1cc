#!/usr/bin/env bash
# One GPU-box visit: A/B sweep of the headline kernel's builds (with the bitwise check), the parity tests, the bench
# lines of both arms, the ncu launch list and one full ncu capture of the headline kernel.
# Everything is written under gpurun_out/ (merged back by gpurun). Each stage has its own timeout so that a slow
# stage cannot eat the visit; stages are ordered by how much the round needs them.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out
mkdir -p $O
date +%s > $O/t_start
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/smi.txt 2>&1

# 1. sweep: which build of the headline kernel is fastest, and are all of them bitwise the general kernel
timeout 300 python scripts/sweep_k1_spec.py --out $O/best_spec.txt > $O/sweep_k1_spec.jsonl 2> $O/sweep.err
echo "sweep rc=$?" >> $O/stages.txt
if [ -s $O/best_spec.txt ]; then export PDEQ_K1_SPEC=$(cat $O/best_spec.txt); fi
echo "PDEQ_K1_SPEC=${PDEQ_K1_SPEC:-unset}" >> $O/stages.txt

# 2. the tests that exercise the changed kernel
timeout 360 python -m pytest tests/test_gpu_k1_spec.py tests/test_gpu_lv_parity.py tests/test_gpu_edge_cases.py -x -q -m gpu > $O/pytest_k1.log 2>&1
echo "pytest_k1 rc=$?" >> $O/stages.txt

# 3. bench lines (ours, then the CPU reference arm)
timeout 240 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" >> $O/stages.txt
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
echo "bench_reference rc=$?" >> $O/stages.txt

# 4. ncu: launch list of the bench command, then one full capture of the headline kernel
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "ncu_launches rc=$?" >> $O/stages.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_loop_kernel --launch-skip 1 -c 1 \
  -o $O/prof_k1_spec -f python scripts/sweep_k1_spec.py --specs ${PDEQ_K1_SPEC:-1} --steps 2 --warmup 1 > $O/ncu_full.log 2>&1
echo "ncu_full rc=$?" >> $O/stages.txt
if [ -f $O/prof_k1_spec.ncu-rep ]; then
  python scripts/summarise_ncu.py $O/prof_k1_spec.ncu-rep $O/k1_spec.ncu.txt > /dev/null 2>> $O/ncu_full.log
fi

# 5. the rest of the GPU suite
timeout ${REST_TIMEOUT:-360} python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_k1_spec.py \
  --ignore tests/test_gpu_lv_parity.py --ignore tests/test_gpu_edge_cases.py --durations=15 > $O/pytest_rest.log 2>&1
echo "pytest_rest rc=$?" >> $O/stages.txt
date +%s > $O/t_end
cat $O/stages.txt
tail -3 $O/pytest_k1.log $O/pytest_rest.log
cat $O/sweep_k1_spec.jsonl | cut -c1-200
