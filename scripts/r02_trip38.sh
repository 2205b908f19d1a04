#!/bin/bash
# Round-2 trip 38: A-collector chunk order as the default: full GPU suite, smoke, bench (chairs full, sintel).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python scripts/parity_probe.py collector 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r02_t38.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t38.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t38.log | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?: $(tail -n 2 gpurun_out/r02_smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('runs_pairs_per_s'))
print('parity', {k: d['parity'].get(k) for k in ('max_abs', 'median_abs', 'frac_px_within_1e-3', 'iterations_within_1e-3', 'timed_path_equals_plain_path')})
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline']['shared_memory']['frac'])
print('corr/lookup', d['roofline_corr_lookup']['frac'], d['roofline_corr_lookup']['ms'])
PY
timeout 400 python bench.py --config sintel --steps 5 --warmup 3 > gpurun_out/r02_bench_sintel.json 2> gpurun_out/r02_bench_sintel.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_sintel.json')); print('sintel', d['value'], 'e2e', d['e2e']['value'], d['parity'].get('max_abs'), d['parity'].get('iterations_within_1e-3'))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches38.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_launches38.log 2>&1; echo "ncu launches exit $?"
