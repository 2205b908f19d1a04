#!/bin/bash
# Round-2 trip 8: resident weights for the 64-channel kRow3 layers (A/B), predict_stream with fixed device buffers.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/parity_probe.py wres 2>&1 | tail -n 1
RAFT_B200_ROW3=2 timeout 300 python scripts/parity_probe.py streamed 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_t8.log 2>&1; echo "pytest exit $? : $(tail -n 1 gpurun_out/r02_t8.log)"; grep -hE "^FAILED|^E  " gpurun_out/r02_t8.log | head -12
timeout 120 python tools/timeline_enc.py 1 8 > gpurun_out/r02_timeline_enc1_wres.log 2>&1; tail -n 5 gpurun_out/r02_timeline_enc1_wres.log
for v in 1 2; do
  RAFT_B200_ROW3=$v timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b8_row3_$v.json 2>gpurun_out/r02_b8_row3_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_b8_row3_$v.json')); print('bench row3=$v', round(d['value'],1), 'pairs/s e2e', round(d['e2e']['value'],1), d['roofline_corr_lookup']['ms'], 'mega ms', round(d['roofline']['ms_per_launch'],4))"
done
timeout 400 python bench.py --config sintel --steps 5 --warmup 3 --quick > gpurun_out/r02_b8_sintel.json 2> gpurun_out/r02_b8_sintel.err; python -c "
import json; d=json.load(open('gpurun_out/r02_b8_sintel.json')); print('sintel', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
