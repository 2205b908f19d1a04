#!/bin/bash
# Round-2 trip 7: correlation store bound fixed; encoder layer timeline; lookup instruction count.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py defaults 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t7.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t7.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t7.log | head -12
for k in 1 6; do timeout 120 python tools/timeline_enc.py $k 8 > gpurun_out/r02_timeline_enc$k.log 2>&1; tail -n 16 gpurun_out/r02_timeline_enc$k.log; done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print('parity', {k: d['parity'].get(k) for k in ('max_abs', 'median_abs', 'frac_px_within_1e-3', 'iterations_within_1e-3', 'timed_path_equals_plain_path')})
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('corr/lookup', d['roofline_corr_lookup']['frac'], d['roofline_corr_lookup']['ms'], d['roofline_corr_lookup']['pyramid_build_alone'], d['roofline_corr_lookup']['lookup_alone_hbm_frac'])
PY
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:corr_lookup_win -s 14 -c 1 python scripts/profile_loop.py f16x2 2 2>&1 | grep -E "inst_executed|duration" 
timeout 400 python bench.py --config sintel --steps 5 --warmup 3 > gpurun_out/r02_bench_sintel.json 2> gpurun_out/r02_bench_sintel.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_sintel.json')); print('sintel', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'].get('max_abs'), d['parity'].get('iterations_within_1e-3'), d['roofline']['achieved'])"
