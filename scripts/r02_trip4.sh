#!/bin/bash
# Round-2 trip 4: pipelined lookup, training-step tests, full bench line, ncu of the new lookup.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t4.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t4.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t4.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?: $(tail -n 2 gpurun_out/r02_smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench exit $?"; tail -n 4 gpurun_out/r02_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print('parity', {k: d['parity'].get(k) for k in ('max_abs', 'median_abs', 'frac_px_within_1e-3', 'iterations_within_1e-3', 'timed_path_equals_plain_path')})
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('corr/lookup', d['roofline_corr_lookup']['frac'], d['roofline_corr_lookup']['ms'], d['roofline_corr_lookup']['pyramid_build_alone'], d['roofline_corr_lookup']['lookup_alone_hbm_frac'])
print('cpu', d['cpu_baseline'])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --quick --sync-e2e > gpurun_out/r02_bench_sync_e2e.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r02_bench_sync_e2e.json')); print('sync e2e', d['value'], d['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_win -s 14 -c 1 -o gpurun_out/r02_prof_lookup3 python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_lookup3.log 2>&1; echo "ncu lookup exit $?"
timeout 300 python tools/corr_sweep.py > gpurun_out/r02_corr_sweep2.log 2>&1; cat gpurun_out/r02_corr_sweep2.log
