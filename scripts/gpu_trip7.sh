#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/t_all.log)"
grep -hE "first sampler|FAILED|Error" gpurun_out/t_all.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f16x2.json 2> gpurun_out/bench_f16x2.err; echo "bench exit $?"; tail -n 2 gpurun_out/bench_f16x2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_f16x2.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['clocks'])
print('  parity', {k: d['parity'][k] for k in ('max_abs','median_abs','frac_px_within_1e-3','iterations_within_1e-3')})
print('  roofline', d['roofline']['achieved'], d['roofline']['executed_frac'], '| corr', d['roofline_corr_lookup']['achieved'], d['roofline_corr_lookup']['ms'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_warm.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_warm.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches_warm.csv --seq 560 16
