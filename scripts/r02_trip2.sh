#!/bin/bash
# Round-2 trip 2: new correlation kernel, window lookup, update_mega_kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run_tests() {  # name, env...
  local name=$1; shift
  env "$@" timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_t2_$name.log 2>&1
  echo "pytest $name exit $? : $(tail -n 1 gpurun_out/r02_t2_$name.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t2_$name.log | head -8
}
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b2_$name.json 2> gpurun_out/r02_b2_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/r02_b2_{name}.json'))
    print(f"bench {name:<14} {d['value']:8.1f} pairs/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  corr/lookup ms {d['roofline_corr_lookup']['ms']} roofline {d['roofline']['achieved']:.1f} TF")
except Exception as e:
    print(f'bench {name}: FAILED ({e})'); print(open(f'gpurun_out/r02_b2_{name}.err').read()[-1500:])
PY
}
run_tests mega A=1
if ! tail -n 1 gpurun_out/r02_t2_mega.log | grep -q " passed" || grep -q failed gpurun_out/r02_t2_mega.log; then run_tests nomega RAFT_B200_MEGA=0; fi
timeout 120 python tools/timeline_corr.py > gpurun_out/r02_timeline_corr.log 2>&1; cat gpurun_out/r02_timeline_corr.log | tail -24
timeout 300 python tools/corr_sweep.py > gpurun_out/r02_corr_sweep.log 2>&1; echo "sweep exit $?"; cat gpurun_out/r02_corr_sweep.log
bench mega A=1
bench nomega RAFT_B200_MEGA=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:update_mega -s 14 -c 1 -o gpurun_out/r02_prof_mega python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_mega.log 2>&1; echo "ncu mega exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_win -s 14 -c 1 -o gpurun_out/r02_prof_lookup python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_lookup.log 2>&1; echo "ncu lookup exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_tc_kernel -s 1 -c 1 -o gpurun_out/r02_prof_corr python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_corr.log 2>&1; echo "ncu corr exit $?"
