#!/bin/bash
# Round-2 trip 31: convf1's im2col rides on the lookup kernel (2 launches per iteration); update-block forms test.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python scripts/parity_probe.py rider 2>&1 | tail -n 3
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_t31.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t31.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t31.log | head -12
for rep in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b31.json 2>gpurun_out/r02_b31.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b31.json')); print('bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), 'mega ms', round(d['roofline']['ms_per_launch'],4), d['roofline_corr_lookup']['ms'], d['gpu_launches'])"
done
