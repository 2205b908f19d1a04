#!/bin/bash
# Round-2 trip 15: update_mega_kernel -- dependency counters polled with nine loads in flight, weights of the first stages
# requested before the dependency wait.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py poll9 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t15.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t15.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t15.log | head -12
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b15.json 2>gpurun_out/r02_b15.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b15.json')); print('bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
timeout 300 python bench.py --config sintel --steps 5 --warmup 3 --quick > gpurun_out/r02_b15s.json 2>gpurun_out/r02_b15s.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b15s.json')); print('sintel', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), 'mega ms', round(d['roofline']['ms_per_launch'],4))"
