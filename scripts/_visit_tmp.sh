set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2i; mkdir -p $O
PDEQ_B200_LIB=$PWD/probdiffeq_b200/lib/legacy/libprobdiffeq_b200_r2f.so timeout 300 python scripts/ab_k2_smoother.py dump /tmp/old.npz > $O/ab_old.log 2>&1
timeout 300 python scripts/ab_k2_smoother.py dump /tmp/new.npz > $O/ab_new.log 2>&1
timeout 120 python scripts/ab_k2_smoother.py compare /tmp/old.npz /tmp/new.npz > $O/ab_compare.jsonl 2>&1
timeout 300 python scripts/bench_configs.py 60 3 8192,65536 > $O/bench_config3.log 2>&1
VISIT=r2i STAGES=ncu_k2 bash scripts/gpu_visit.sh > /dev/null 2>&1
cat $O/ab_compare.jsonl; tail -3 $O/ab_new.log; cat $O/bench_config3.log; head -48 $O/k2.ncu.txt
