#!/bin/bash
# encoder investigation: timelines of encoder convolutions, promotion-group size A/B
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in 0 1 6; do timeout 120 python tools/timeline_enc.py $k > gpurun_out/timeline_enc$k.log 2>&1; tail -n 16 gpurun_out/timeline_enc$k.log; done
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench($name) exit $?"
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_$name.json'))
print('$name', {k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
PY
}
run default A=1
run grp3 RAFT_B200_ENC_GROUP=3
run grp5 RAFT_B200_ENC_GROUP=5
RAFT_B200_ENC_GROUP=5 timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "encoder or 448" > gpurun_out/t_grp5.log 2>&1
echo "pytest(grp5) exit $? : $(tail -n 1 gpurun_out/t_grp5.log)"; grep -hE "FAILED|max-abs|assert" gpurun_out/t_grp5.log | head -8
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "pytest exit $? : $(tail -n 1 gpurun_out/t_all.log)"
