#!/bin/bash
# A/B of update_mega_kernel producer variants on ONE box: builds with RAFT_MEGA_POLL9 / RAFT_MEGA_PREB = 0/1, bench --quick
# of each, twice, interleaved.     build (no GPU): bash tools/mega_ab.sh build      run: bash tools/mega_ab.sh run
set -u
cd "$(dirname "$0")/.."
mkdir -p tools/epi_exp
if [ "${1:-run}" = build ]; then
  for v in 00 10 01 11; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
      -DRAFT_MEGA_POLL9=${v:0:1} -DRAFT_MEGA_PREB=${v:1:1} tf_raft_b200/csrc/api.cu -o tools/epi_exp/mega_$v.so &
  done
  wait
  ls -la tools/epi_exp/mega_*.so
  exit 0
fi
for rep in 1 2; do
  for v in 00 10 01 11; do
    RAFT_B200_LIB=$PWD/tools/epi_exp/mega_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --quick > /tmp/ab_$v.json 2>/dev/null
    python -c "
import json; d=json.load(open('/tmp/ab_$v.json')); print('poll9/preb $v rep $rep:', round(d['value'],1), 'pairs/s  mega ms', round(d['roofline']['ms_per_launch'],4))"
  done
done
