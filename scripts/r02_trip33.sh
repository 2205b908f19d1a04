#!/bin/bash
# Round-2 trip 33: the driver's commands on the final tree (tests, smoke, bench, reference arm).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_t33.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t33.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?: $(tail -n 2 gpurun_out/r02_smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('runs_pairs_per_s'), d['clocks'])
print('parity', d['parity'].get('max_abs'), d['parity'].get('timed_path_equals_plain_path'))
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('corr/lookup', d['roofline_corr_lookup']['frac'], d['roofline_corr_lookup']['ms'], d['roofline_corr_lookup']['lookup_alone_hbm_frac'])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "reference exit $?"; head -c 200 gpurun_out/r02_bench_reference.json; echo
timeout 300 python bench.py > gpurun_out/r02_bench_default.json 2>/dev/null; echo "default-flags bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_default.json')); print(d['value'], d['steps'], d['warmup'], d['e2e']['value'])"
