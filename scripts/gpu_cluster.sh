#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for CL in 2 4; do
  export RAFT_B200_CLUSTER=$CL
  timeout 600 python -m pytest tests -m gpu -q -x -k "f16x2 or stages or api_contract or properties" -p no:cacheprovider > gpurun_out/t_cl$CL.log 2>&1; echo "CL=$CL pytest exit $? : $(tail -n 1 gpurun_out/t_cl$CL.log)"
  grep -hE "FAILED|Error|error" gpurun_out/t_cl$CL.log | head -5
done
for CL in 1 2 4; do
  export RAFT_B200_CLUSTER=$CL
  timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cl$CL.json 2> gpurun_out/bench_cl$CL.err; echo "CL=$CL bench exit $?"
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_cl$CL.json'))
print('CL=$CL', {k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'max_abs', d['parity']['max_abs'], 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['executed_frac'],3), '| corr', round(d['roofline_corr_lookup']['achieved']), d['roofline_corr_lookup']['ms'])
PY
done
export RAFT_B200_CLUSTER=4
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_cl4.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/ncu_cl4.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches_cl4.csv --seq 560 15 | head -8; python scripts/ncu_summary.py gpurun_out/launches_cl4.csv --seq 560 15 | tail -15
