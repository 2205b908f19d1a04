#!/bin/bash
# Round-2 trip 5: promoted correlation (parity), kRow3 encoder layers (descriptor rule A/B), lookup v4, train tests.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
probe() { local tag=$1; shift; env "$@" timeout 300 python scripts/parity_probe.py $tag 2>&1 | tail -n 1; }
probe row3_off RAFT_B200_ROW3=0
probe row3_doc RAFT_B200_ROW3=1
probe row3_nobase RAFT_B200_ROW3=2
probe row3off_nomega RAFT_B200_ROW3=0 RAFT_B200_MEGA=0
RAFT_B200_ROW3=0 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t5_row3off.log 2>&1; echo "pytest row3=0 exit $? : $(tail -n 1 gpurun_out/r02_t5_row3off.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t5_row3off.log | head -12
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "encoder or 448 or reference_test_shape or other_resolutions or graph" > gpurun_out/r02_t5_row3.log 2>&1; echo "pytest row3=1 exit $? : $(tail -n 1 gpurun_out/r02_t5_row3.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t5_row3.log | head -12
RAFT_B200_ROW3=2 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "encoder or 448 or reference_test_shape or other_resolutions or graph" > gpurun_out/r02_t5_row3b.log 2>&1; echo "pytest row3=2 exit $? : $(tail -n 1 gpurun_out/r02_t5_row3b.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t5_row3b.log | head -6
for v in 0 1 2; do
  RAFT_B200_ROW3=$v timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b5_row3_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02_b5_row3_$v.json')); print('bench row3=$v', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
done
