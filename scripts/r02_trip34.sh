#!/bin/bash
# Round-2 trip 34: bench.py after the shared-memory roofline object was added (chairs, sintel, train; quick).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cfg in chairs sintel train; do
timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --quick > gpurun_out/r02_b34_$cfg.json 2>gpurun_out/r02_b34_$cfg.err; echo "$cfg exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_b34_$cfg.json')); print('$cfg', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('roofline',{}).get('shared_memory'))" || tail -n 5 gpurun_out/r02_b34_$cfg.err
done
