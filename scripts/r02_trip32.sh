#!/bin/bash
# Round-2 trip 32: final evidence run (lookup kernel carries the im2col) (tests, smoke, bench, reference arm, launch list, ncu of the mega kernel, sintel, train).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py final 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t32.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t32.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t32.log | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?: $(tail -n 2 gpurun_out/r02_smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print('parity', {k: d['parity'].get(k) for k in ('max_abs', 'median_abs', 'frac_px_within_1e-3', 'iterations_within_1e-3', 'timed_path_equals_plain_path')})
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('corr/lookup', d['roofline_corr_lookup']['frac'], d['roofline_corr_lookup']['ms'], d['roofline_corr_lookup']['pyramid_build_alone'], d['roofline_corr_lookup']['lookup_alone_hbm_frac'])
print('cpu', d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_reference.json')); print('reference arm', d['value'], d['cpu_baseline'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches32.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_launches32.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:update_mega -s 14 -c 1 -o gpurun_out/r02_prof_mega5 python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_mega5.log 2>&1; echo "ncu mega exit $?"
timeout 400 python bench.py --config sintel --steps 5 --warmup 3 > gpurun_out/r02_bench_sintel.json 2> gpurun_out/r02_bench_sintel.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_sintel.json')); print('sintel', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity'].get('max_abs'), d['parity'].get('iterations_within_1e-3'), d['roofline']['achieved'])"
timeout 400 python bench.py --config train --steps 5 --warmup 3 > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_train.json')); print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['loss_first_last'], d['cpu_baseline'])"
