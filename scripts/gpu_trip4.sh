#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -s -k "stages or f16x2" -p no:cacheprovider > gpurun_out/t_tc.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/t_tc.log)"
grep -hE "max-abs|FAILED|passed|failed" gpurun_out/t_tc.log | head -40
timeout 300 python scripts/diag_e2e.py > gpurun_out/diag_e2e.log 2>&1; echo "diag exit $?"
grep -A 18 "f16x2: manual" gpurun_out/diag_e2e.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:raft --csv --log-file gpurun_out/launches_loop.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_loop.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches_loop.csv --seq 190 30
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f16x2.json 2> gpurun_out/bench_f16x2.err; echo "bench exit $?"
python -c "
import json; d = json.load(open('gpurun_out/bench_f16x2.json')); print({k: d[k] for k in ('value','ms_per_step','final_flow_max_abs_vs_oracle','gpu_launches')}); print(d['roofline']); print(d['roofline_corr_lookup'])"
