#!/bin/bash
# Round-2 trip 12: one fence per block in norm_stats_kernel (trip 11 measured +52 us per launch for the fused finalisation);
# 32-byte loads of the residual / GRU-h operands in the epilogues.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py ldg256 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t12.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t12.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t12.log | head -12
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b12.json 2>gpurun_out/r02_b12.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b12.json')); print('bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches12.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_launches12.log 2>&1; echo "ncu launches exit $?"
python scripts/ncu_summary.py gpurun_out/r02_launches12.csv | head -12
