#!/bin/bash
# Round-2 trip 37: A-collector orders of the three-pass chunk (tools/epi_exp/coll_<order>.so): parity of the benchmark pair,
# update-block and encoder tests, bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 0 1 2 3; do
  export RAFT_B200_LIB=$PWD/tools/epi_exp/coll_$v.so
  timeout 200 python scripts/parity_probe.py order$v 2>&1 | tail -n 1
  timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r02_b37_$v.json 2>gpurun_out/r02_b37_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_b37_$v.json')); print('order $v bench', round(d['value'],1), 'pairs/s  mega ms', round(d['roofline']['ms_per_launch'],4), d['roofline_corr_lookup']['ms'])"
done
