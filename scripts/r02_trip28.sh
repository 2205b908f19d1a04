#!/bin/bash
# Round-2 trip 28: two MMA-issuing warps (alternate promotion groups).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python scripts/parity_probe.py issuer2 2>&1 | tail -n 3
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_t28.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t28.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t28.log | head -12
for rep in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b28.json 2>gpurun_out/r02_b28.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b28.json')); print('bench', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), 'mega ms', round(d['roofline']['ms_per_launch'],4))"
done
RAFT_B200_LIB=$PWD/tools/epi_exp/mega_trace.so timeout 200 python tools/timeline_mega.py > gpurun_out/r02_timeline_mega_two_issuers.log 2>&1; cat gpurun_out/r02_timeline_mega_two_issuers.log | tail -n 20
