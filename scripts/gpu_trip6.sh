#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/t_all.log)"
grep -hE "first sampler|FAILED|Error" gpurun_out/t_all.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f16x2.json 2> gpurun_out/bench_f16x2.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench_f16x2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-graph > gpurun_out/bench_f16x2_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph exit $?"
python - <<'PY'
import json
for f in ('bench_f16x2', 'bench_f16x2_nograph'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
        print('  parity', {k: d['parity'][k] for k in ('max_abs','median_abs','frac_px_within_1e-3','iterations_within_1e-3')})
        print('  roofline', d['roofline']['achieved'], d['roofline']['executed_frac'], '| corr', d['roofline_corr_lookup']['achieved'], d['roofline_corr_lookup']['ms'])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_warm.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_warm.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches_warm.csv --seq 560 16
