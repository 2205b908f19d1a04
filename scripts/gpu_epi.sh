#!/bin/bash
# epilogue experiment: parity tests + bench + timeline at 16 and 8 promotion/epilogue warps
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "pytest exit $? : $(tail -n 1 gpurun_out/t_all.log)"
grep -hE "FAILED|Error" gpurun_out/t_all.log | head -10
for w in 16 8; do
  export RAFT_B200_EPI_WARPS=$w
  timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_epi$w.json 2> gpurun_out/bench_epi$w.err; echo "bench($w) exit $?"
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_epi$w.json'))
print('epi warps $w', {k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'parity', d['parity']['max_abs'], d['parity']['iterations_within_1e-3'])
PY
  for l in 4 0 5 8; do timeout 120 python tools/timeline.py $l 2>&1 | tail -n 1; done
done
