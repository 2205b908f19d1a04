#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -s -k "encoder" -p no:cacheprovider > gpurun_out/t_enc.log 2>&1; echo "enc pytest exit $? : $(tail -n 1 gpurun_out/t_enc.log)"
grep -hE "max-abs|FAILED|Error|error" gpurun_out/t_enc.log | head -30
timeout 900 python -m pytest tests -m gpu -q -k "not encoder" -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all pytest exit $? : $(tail -n 1 gpurun_out/t_all.log)"
grep -hE "FAILED|Error" gpurun_out/t_all.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f16x2.json 2> gpurun_out/bench_f16x2.err; echo "bench exit $?"
python -c "
import json; d = json.load(open('gpurun_out/bench_f16x2.json')); print({k: d[k] for k in ('value','ms_per_step','final_flow_max_abs_vs_oracle','gpu_launches')}); print(d['e2e']); print(d['roofline']); print(d['roofline_corr_lookup']); print(d['cpu_baseline'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_fwd.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches_fwd.csv
