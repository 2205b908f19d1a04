#!/bin/bash
# Round-2 trip 14: update_mega_kernel claims its items dynamically (atomic cursor) and the flow branch is interleaved with the
# correlation branch in the list; norm finalisation back in its own kernel; packed fp16 conversions in the re-split.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py dynclaim 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t14.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t14.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t14.log | head -12
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b14.json 2>gpurun_out/r02_b14.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b14.json')); print('bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches14.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_launches14.log 2>&1; echo "ncu launches exit $?"
python scripts/ncu_summary.py gpurun_out/r02_launches14.csv 2>/dev/null | head -12
