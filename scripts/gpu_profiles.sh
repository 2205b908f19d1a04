#!/bin/bash
# Evidence run for profiles/: tests, bench, ncu launch list (recipe flags), ncu --set full on the dominant kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/t_all.log)"
grep -hE "first sampler|FAILED|Error" gpurun_out/t_all.log | head -10
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?: $(tail -n 2 gpurun_out/smoke.log | tr '\n' ' ')"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench exit $?"
kill $SMI
if [ -z "${FAST:-}" ]; then
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32 > gpurun_out/bench_r01_fp32.json 2> gpurun_out/bench_r01_fp32.err; echo "bench fp32 exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> /dev/null; echo "bench ref exit $?"
fi
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 232 -c 3 -o gpurun_out/prof_r01_conv_tc python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 300 ncu --set full --clock-control none -k regex:corr_lookup_kernel -s 14 -c 1 -o gpurun_out/prof_r01_lookup python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_full2.log 2>&1; echo "ncu lookup exit $?"
python tools/timeline.py 4 > gpurun_out/timeline_zr1.log 2>&1
[ -z "${FAST:-}" ] && ./tools/tma_probe > gpurun_out/tma_probe.log 2>&1
[ -z "${FAST:-}" ] && timeout 300 python tools/corr_sweep.py > gpurun_out/corr_sweep.log 2>&1
for l in 0 5 8; do python tools/timeline.py $l 2>&1 | tail -n 1 >> gpurun_out/timeline_zr1.log; done
python tools/timeline_enc.py 1 > gpurun_out/timeline_enc1.log 2>&1
RAFT_B200_ENC_GROUP=2 timeout 200 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/bench_r01_encgroup2.json 2> /dev/null; echo "bench encgroup2 exit $?"
RAFT_B200_UPD_GROUP=3 timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_updgroup3.json 2> /dev/null; echo "bench updgroup3 exit $?"
python - <<'PY'
import json
for f in ('bench_r01', 'bench_r01_updgroup3', 'bench_r01_encgroup2'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k: d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
        if 'max_abs' in d.get('parity', {}): print('  parity', {k: d['parity'][k] for k in ('max_abs','median_abs','frac_px_within_1e-3','iterations_within_1e-3')}, d['clocks'])
        if 'roofline' in d and d.get('cpu_baseline', {}).get('value'): print('  roofline', d['roofline']['achieved'], d['roofline']['executed_frac'], '| corr', d['roofline_corr_lookup']['achieved'], d['roofline_corr_lookup']['ms'], '| cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
    except Exception as e: print(f, 'ERR', e)
PY
