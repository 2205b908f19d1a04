#!/bin/bash
# Cost breakdown of the convolution epilogues: builds of the library with parts of the epilogue compiled out
# (RAFT_EPI_EXP bit 0: no 32-byte stores, 1: no operand loads, 2: no MUFU in the gates, 3: no fp16 conversions)
# and the per-tile timeline of three update-block layers (1 = convc2, 4 = GRU zr, 8 = flow-head conv1) for each.
#   build (no GPU):  bash tools/epi_exp.sh build        run (GPU box):  bash tools/epi_exp.sh run
set -u
cd "$(dirname "$0")/.."
VARIANTS="0 1 2 4 8 15"
if [ "${1:-run}" = build ]; then
  for v in $VARIANTS; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DRAFT_EPI_EXP=$v \
      tf_raft_b200/csrc/api.cu -o tools/epi_exp/lib_$v.so &
  done
  wait
  ls -la tools/epi_exp/
  exit 0
fi
for v in $VARIANTS; do
  for layer in 1 4 8; do
    echo "variant $v layer $layer: $(RAFT_B200_MEGA=0 RAFT_B200_LIB=$PWD/tools/epi_exp/lib_$v.so timeout 120 python tools/timeline.py $layer 2>&1 | tail -n 1)"
  done
done
