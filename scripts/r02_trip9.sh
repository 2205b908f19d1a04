#!/bin/bash
# Round-2 trip 9: L2 prefetch of the next tile's activation boxes in the encoder convolutions (A/B), e2e pipeline.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py prefetch 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t9.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t9.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t9.log | head -12
for k in 1 6; do timeout 120 python tools/timeline_enc.py $k 8 > gpurun_out/r02_timeline_enc${k}_pf.log 2>&1; tail -n 4 gpurun_out/r02_timeline_enc${k}_pf.log; done
for v in 1 0; do
  RAFT_B200_ENC_PREFETCH=$v timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b9_pf_$v.json 2>gpurun_out/r02_b9_pf_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_b9_pf_$v.json')); print('bench prefetch=$v', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches9.csv python scripts/profile_loop.py f16x2 2 > gpurun_out/r02_ncu_launches9.log 2>&1; echo "ncu launches exit $?"
