#!/bin/bash
# Round-2 trip 10: TMA probe of the encoder box shapes; 32-byte epilogue stores.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 ./tools/tma_probe2 > gpurun_out/r02_tma_probe2.log 2>&1; cat gpurun_out/r02_tma_probe2.log
timeout 300 python scripts/parity_probe.py st256 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t10.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t10.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t10.log | head -12
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b10.json 2>gpurun_out/r02_b10.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b10.json')); print('bench st256', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
timeout 120 python tools/timeline_enc.py 1 8 > gpurun_out/r02_timeline_enc1_st256.log 2>&1; tail -n 4 gpurun_out/r02_timeline_enc1_st256.log
